// A C++ caller written against the REFERENCE's class interface (src/dft_solver.h:7-87): it uses only what that header
// declares -- the three solver classes through values, through a base pointer (virtual destructor), the protected
// safe_cublas_dgemm from a subclass, compute_coulomb, and the four extern "C" functions.  tests/test_cxx_client.py
// compiles it with g++ against include/dft_solver.h AND (where staged) against the reference's own header, links it
// with this repo's weights/dft.so, and on a GPU runs it and checks the numbers against the oracle.  Test scaffolding.
//
//   cxx_client <input.bin> <output.bin>
//   input : int32 ngrid, nao | dm (nao,nao) | ao (ngrid,nao) | grad (3,ngrid,nao) | w (ngrid) | eri (nao^2,nao^2)
//   output: for type 0,1,2: E (class API), V (nao,nao) | 3 x E (C API through DFT_ComputeXC) | J class, J C API (nao,nao
//           each) | C = A^T B (3 x 2) from safe_cublas_dgemm with A (4 x 3), B (4 x 2) column-major
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include "dft_solver.h"

namespace {

struct Dev {
    double* p = nullptr;
    explicit Dev(size_t n, const double* h = nullptr) {
        if (cudaMalloc(reinterpret_cast<void**>(&p), (n ? n : 1) * sizeof(double)) != cudaSuccess) { fprintf(stderr, "cudaMalloc failed\n"); exit(3); }
        if (h) cudaMemcpy(p, h, n * sizeof(double), cudaMemcpyHostToDevice);
        else cudaMemset(p, 0, (n ? n : 1) * sizeof(double));
    }
    ~Dev() { cudaFree(p); }
    std::vector<double> get(size_t n) const {
        std::vector<double> h(n);
        cudaMemcpy(h.data(), p, n * sizeof(double), cudaMemcpyDeviceToHost);
        return h;
    }
    unsigned long long u() const { return reinterpret_cast<unsigned long long>(p); }
};

// a derived class may call the protected GEMM helper (dft_solver.h:25-27)
struct GemmProbe : public LDASolver {
    void atb(int m, int n, int k, const double* A, int lda, const double* B, int ldb, double* C, int ldc) {
        safe_cublas_dgemm(true, false, m, n, k, A, lda, B, ldb, C, ldc);
    }
};

std::vector<double> read_doubles(FILE* f, size_t n) {
    std::vector<double> v(n);
    if (n && fread(v.data(), sizeof(double), n, f) != n) { fprintf(stderr, "short input\n"); exit(2); }
    return v;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc != 3) return 1;
    FILE* in = fopen(argv[1], "rb");
    if (!in) return 2;
    int dims[2];
    if (fread(dims, sizeof(int), 2, in) != 2) return 2;
    const size_t ngrid = dims[0], nao = dims[1], n2 = nao * nao;
    auto dm = read_doubles(in, n2), ao = read_doubles(in, ngrid * nao), grad = read_doubles(in, 3 * ngrid * nao),
         w = read_doubles(in, ngrid), eri = read_doubles(in, n2 * n2);
    fclose(in);
    Dev d_dm(n2, dm.data()), d_ao(ngrid * nao, ao.data()), d_grad(3 * ngrid * nao, grad.data()), d_w(ngrid, w.data()),
        d_eri(n2 * n2, eri.data()), d_v(n2), d_J(n2);
    std::vector<double> out;

    // 1. the classes, as values and through the abstract base (virtual compute_xc, virtual destructor)
    {
        LDASolver lda;
        out.push_back(lda.compute_xc((int)ngrid, (int)nao, d_dm.p, d_ao.p, nullptr, d_w.p, d_v.p));
        auto v = d_v.get(n2);
        out.insert(out.end(), v.begin(), v.end());
        std::unique_ptr<XCSolver> gga(new GGASolver());
        out.push_back(gga->compute_xc((int)ngrid, (int)nao, d_dm.p, d_ao.p, d_grad.p, d_w.p, d_v.p));
        v = d_v.get(n2);
        out.insert(out.end(), v.begin(), v.end());
        XCSolver* b3 = new B3LYPSolver();
        out.push_back(b3->compute_xc((int)ngrid, (int)nao, d_dm.p, d_ao.p, d_grad.p, d_w.p, d_v.p));
        v = d_v.get(n2);
        out.insert(out.end(), v.begin(), v.end());
        delete b3;
    }
    // 2. the C ABI
    std::vector<double> j_c;
    for (int type = SOLVER_LDA; type <= SOLVER_B3LYP; ++type) {
        XCSolver* s = DFT_CreateSolver(type);
        if (!s) { fprintf(stderr, "DFT_CreateSolver(%d) returned NULL\n", type); return 4; }
        out.push_back(DFT_ComputeXC(s, (int)ngrid, (int)nao, d_dm.u(), d_ao.u(), type ? d_grad.u() : 0ull, d_w.u(), d_v.u()));
        if (type == SOLVER_B3LYP) {
            DFT_ComputeCoulomb(s, (int)nao, d_eri.u(), d_dm.u(), d_J.u());
            cudaDeviceSynchronize();
            j_c = d_J.get(n2);
        }
        DFT_DestroySolver(s);
    }
    if (DFT_CreateSolver(7) != nullptr) return 5;                     // unknown type -> nullptr (dft_solver.cu:681)
    if (DFT_ComputeXC(nullptr, 1, 1, 0, 0, 0, 0, 0) != 0.0) return 6; // null solver -> 0.0 (dft_solver.cu:695)
    // 3. compute_coulomb through the class, the protected GEMM through a subclass
    {
        GemmProbe probe;
        probe.compute_coulomb((int)nao, d_eri.p, d_dm.p, d_J.p);
        cudaDeviceSynchronize();
        auto j = d_J.get(n2);
        out.insert(out.end(), j.begin(), j.end());
        out.insert(out.end(), j_c.begin(), j_c.end());
        double A[12], B[8];
        for (int i = 0; i < 12; ++i) A[i] = 0.25 * i - 1.0;           // (4 x 3), column-major, lda 4
        for (int i = 0; i < 8; ++i) B[i] = 1.0 / (1.0 + i);           // (4 x 2), column-major, ldb 4
        Dev dA(12, A), dB(8, B), dC(6);
        probe.atb(3, 2, 4, dA.p, 4, dB.p, 4, dC.p, 3);
        cudaDeviceSynchronize();
        auto c = dC.get(6);
        out.insert(out.end(), c.begin(), c.end());
    }
    if (cudaDeviceSynchronize() != cudaSuccess) return 7;
    FILE* o = fopen(argv[2], "wb");
    if (!o) return 8;
    fwrite(out.data(), sizeof(double), out.size(), o);
    fclose(o);
    return 0;
}
