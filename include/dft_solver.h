// dft_solver.h -- the drop-in boundary of the B200-native XC engine.
//
// This header declares exactly the interface the reference declares in
// /root/reference/src/dft_solver.h, so that a library built from
// quantum_compute_dft_b200/csrc/ can replace the reference's weights/dft.so
// under its unmodified Python driver (dft.py:24-50 binds the four C symbols
// with ctypes).  Each declaration cites the reference line it replaces.
//
//   class XCSolver                      dft_solver.h:7-30   (abstract base, non-copyable,
//                                                            virtual dtor, pure-virtual compute_xc,
//                                                            public compute_coulomb, protected
//                                                            safe_cublas_dgemm, unique_ptr pimpl)
//   class LDASolver / GGASolver /       dft_solver.h:32-63  (ctor + compute_xc override each)
//         B3LYPSolver
//   enum SolverType                     dft_solver.h:67-71
//   DFT_CreateSolver                    dft_solver.h:73     body dft_solver.cu:677-682
//   DFT_DestroySolver                   dft_solver.h:75     body dft_solver.cu:684-686
//   DFT_ComputeXC                       dft_solver.h:77-82  body dft_solver.cu:688-704
//   DFT_ComputeCoulomb                  dft_solver.h:84-87  body dft_solver.cu:706-718
//
// `CublasHandleWrapper` is only forward-declared by the reference (dft_solver.h:5);
// here its definition (csrc/engine.h) holds the engine context -- stream, grow-only
// workspace, TMA descriptors, communicator -- and no cuBLAS handle at all.
//
// Semantics kept from the reference:
//   * all pointers are DEVICE pointers to C-contiguous float64 arrays on the current device:
//       d_dm (nao,nao)  d_ao (ngrid,nao)  d_ao_grad (3,ngrid,nao) or NULL for LDA
//       d_weights (ngrid)  d_vxc (nao,nao) overwritten   d_eri (nao^2,nao^2)   d_J (nao,nao)
//   * compute_xc returns E_xc = sum_g w_g rho_g eps_xc(rho_g, sigma_g) as a host double and has
//     finished all device work on return (the reference blocks on an 8-byte D2H, dft_solver.cu:575);
//   * the consumer forms 1/2 (V + V^T) from d_vxc (dft.py:212).  The reference leaves a
//     different raw matrix per functional (symmetric for LDA, unsymmetrised B^T Phi for GGA,
//     M + M^T for B3LYP); this engine always writes the symmetric matrix S with
//     1/2 (S + S^T) == 1/2 (V_ref + V_ref^T), so the driver's result is unchanged;
//   * no exceptions and no error codes cross the C ABI; a null solver is a no-op that
//     returns 0.0 (dft_solver.cu:695,711); an unknown type gives nullptr (:681).
//     Deviation (documented in INTEGRATION.md): after a failed CUDA call the engine prints
//     the error to stderr like the reference does, and DFT_ComputeXC returns NaN instead of
//     a number computed from garbage.
//
// Additive entry points (AO evaluation on the GPU, multi-GPU, options) are declared in
// dft_b200_ext.h; none of them changes the four symbols above.
#pragma once
#include <memory>
#include <vector>

struct CublasHandleWrapper;  // engine context (pimpl); defined in csrc/engine.h

class XCSolver {
public:
    XCSolver();
    virtual ~XCSolver();
    XCSolver(const XCSolver&) = delete;
    XCSolver& operator=(const XCSolver&) = delete;

    // E_xc returned; V_xc written to d_vxc.  Reference: dft_solver.h:15-20.
    virtual double compute_xc(int ngrid, int nao,
                              const double* d_dm, const double* d_ao, const double* d_ao_grad,
                              const double* d_weights, double* d_vxc) = 0;

    // J_ij = sum_kl (ij|kl) D_kl over the dense ERI.  Reference: dft_solver.h:22, dft_solver.cu:550-555.
    void compute_coulomb(int nao, const double* d_eri, const double* d_dm, double* d_J);

    // Engine context accessor for the additive C entry points (not in the reference).
    CublasHandleWrapper* context() { return handle_wrapper.get(); }

protected:
    // Kept so that code written against the reference header still links
    // (dft_solver.h:25-27).  Implemented by a plain 16 x 16-tile SIMT FP64 GEMM (csrc/linalg.cu,
    // gemm_simple_kernel): correct for any shape and transposition, NOT tuned and NOT on the XC hot
    // path -- the engine's contractions are the fused DMMA kernels in csrc/xc_tma.cu.
    void safe_cublas_dgemm(bool transA, bool transB, int m, int n, int k,
                           const double* A, int lda, const double* B, int ldb,
                           double* C, int ldc);

    std::unique_ptr<CublasHandleWrapper> handle_wrapper;
};

class LDASolver : public XCSolver {  // Slater exchange + VWN5 correlation
public:
    LDASolver();
    double compute_xc(int ngrid, int nao, const double* d_dm, const double* d_ao,
                      const double* d_ao_grad, const double* d_weights, double* d_vxc) override;
};

class GGASolver : public XCSolver {  // PBE exchange + PBE correlation
public:
    GGASolver();
    double compute_xc(int ngrid, int nao, const double* d_dm, const double* d_ao,
                      const double* d_ao_grad, const double* d_weights, double* d_vxc) override;
};

class B3LYPSolver : public XCSolver {  // 0.80 Slater + 0.72 dB88 + 0.19 VWN-RPA + 0.81 LYP (local part)
public:
    B3LYPSolver();
    double compute_xc(int ngrid, int nao, const double* d_dm, const double* d_ao,
                      const double* d_ao_grad, const double* d_weights, double* d_vxc) override;
};

extern "C" {
enum SolverType { SOLVER_LDA = 0, SOLVER_GGA = 1, SOLVER_B3LYP = 2 };

XCSolver* DFT_CreateSolver(int type);
void DFT_DestroySolver(XCSolver* solver);

double DFT_ComputeXC(XCSolver* solver, int ngrid, int nao,
                     unsigned long long d_dm_ptr, unsigned long long d_ao_ptr,
                     unsigned long long d_ao_grad_ptr, unsigned long long d_weights_ptr,
                     unsigned long long d_vxc_ptr);

void DFT_ComputeCoulomb(XCSolver* solver, int nao,
                        unsigned long long d_eri_ptr, unsigned long long d_dm_ptr,
                        unsigned long long d_J_ptr);
}
