// dft_b200_ext.h -- ADDITIVE C entry points of the B200-native XC engine.
//
// None of these exists in the reference; none changes the four symbols of
// dft_solver.h.  They cover what BASELINE.json's north_star adds around the
// drop-in path:
//   (a) AO evaluation on the GPU  -> replaces numint.eval_ao at grid.py:30,38 and the
//       H2D copies of the AO arrays at dft.py:155,172;
//   (e) grid-sharded multi-GPU    -> one process per GPU, NCCL all-reduce of the
//       nao x nao partial V_xc and the scalar E_xc inside DFT_ComputeXC;
//   options / statistics / microbenchmarks used by bench.py for the roofline denominators.
// All functions return 0 on success and a non-zero code on failure unless stated otherwise;
// like the reference ABI they never throw.
#pragma once
#include "dft_solver.h"

extern "C" {

// ---- (a) AO evaluation --------------------------------------------------------------------
// Values (deriv=0) or values + first derivatives (deriv=1) of contracted Cartesian s/p Gaussian
// shells on ngrid points.  Writes exactly the layouts DFT_ComputeXC consumes:
//   d_ao      (ngrid, nao)    row-major float64          (dft.py:155)
//   d_ao_grad (3, ngrid, nao) planar x,y,z, or 0 if deriv=0 (dft.py:136-142,172)
// d_coords is a DEVICE pointer to (ngrid,3) float64 Bohr coordinates.  The shell tables are HOST
// arrays (they are tiny and are staged into shared memory by the kernel):
//   shell_xyz (nshell,3), shell_l (0|1), shell_ao_off, shell_prim_off, shell_nprim,
//   prim_exp / prim_coef (nprim_total; coefficient already includes the primitive norm).
// A primitive contributes only where exp*r^2 <= exp_cutoff (PySCF-like screening); pass <= 0
// for the default (60).
int DFT_EvalAO(XCSolver* solver, int ngrid, unsigned long long d_coords_ptr,
               int nshell, const double* shell_xyz, const int* shell_l,
               const int* shell_ao_off, const int* shell_prim_off, const int* shell_nprim,
               int nprim_total, const double* prim_exp, const double* prim_coef,
               int nao, int deriv, double exp_cutoff,
               unsigned long long d_ao_ptr, unsigned long long d_ao_grad_ptr);

// ---- (e) multi-GPU: one process per GPU ---------------------------------------------------
// Rank 0 calls DFT_CommGetUniqueId and ships the 128 bytes to the other ranks out of band
// (bench.py uses a torch.distributed / TCP store broadcast); every rank then calls
// DFT_CommInit.  From then on DFT_ComputeXC treats its (ngrid, d_ao, d_ao_grad, d_weights)
// as THIS RANK's slice of the grid, all-reduces (sum) the nao*nao partial V_xc and E_xc over
// NVLink, and every rank returns the global E_xc and holds the global V_xc.  NCCL is loaded
// with dlopen("libnccl.so.2") on first use; a single-GPU caller never needs it.
int DFT_CommGetUniqueId(void* out_id_128_bytes);
int DFT_CommInit(XCSolver* solver, int rank, int nranks, const void* id_128_bytes);
int DFT_CommDestroy(XCSolver* solver);

// ---- (e) multi-GPU: ONE process, behind the unmodified DFT_ComputeXC (SURVEY.md 8e "process model") ----------
// The reference's driver holds every array on one device (dft.py:155-176) and calls DFT_ComputeXC from one process.
// DFT_SetOption(solver, "devices", n) -- or the environment variable DFT_B200_DEVICES=n|all, read by DFT_CreateSolver,
// for a driver that knows nothing of options -- makes that same call use n GPUs of the box (csrc/fanout.cu): the grid is
// dealt to one child engine per device in interleaved blocks of 1024 points, each child keeps its shard of
// (ao, ao_grad, weights) resident on its own device (pulled out of the caller's arrays over NVLink by the copy engines
// ONCE and reused while the caller passes the same arrays: same pointers, sizes and content fingerprint -- all weights
// and 2^17 samples per plane), and per call only D (nao^2 doubles) goes out and [V_xc | E_xc] comes back, summed in a
// fixed order by one kernel on the caller's device through peer loads.  d_vxc, the returned E_xc and the blocking
// semantics are exactly those of the single-GPU call.
//   option keys: "devices" n (n <= visible devices; 1 tears the fan-out down), "virtual_devices" n (n children dealt
//       round-robin over the visible devices: the whole mechanism on a one-GPU box, for tests),
//       "devices_min_work" x (builds with ngrid * nao^2 < x stay on the caller's device, default 2e9: H2O, benzene),
//       "ao_cache" 0|1 (default 1; 0 re-cuts the shards on every call), "ao_invalidate" (drop the resident shards now:
//       for a caller that rewrites its AO arrays in place in a way 2^17 samples per plane might miss).
//       "fan_threads" 0|1 (default 1: a worker thread per device enqueues that device's share of a call; 0: the caller's
//       thread does it for all devices in turn -- at 8 devices the last one then starts 0.3 ms after the first).
//   environment (read in DFT_CreateSolver, for a driver that cannot set options): DFT_B200_DEVICES=n|all,
//       DFT_B200_VIRTUAL_DEVICES=n, DFT_B200_DEVICES_MIN_WORK=x; DFT_B200_VERBOSE=1 reports every shard cut on stderr.
//   stat keys: "devices", "fan_active" (the last call was fanned out), "fan_scatters" (times the shards were cut),
//       "fan_peer_loads" (the reduction reads the children's results through peer loads), "fan_resident_bytes";
//       "density_ms" / "vxc_ms" / "reduce_ms" are those of the slowest device, "total_ms" the whole call.
// Points of shard `shard` of that deal (host arithmetic only; -1 on bad arguments).
int DFT_ShardPoints(int ngrid, int nshards, int shard);

// ---- options / statistics ------------------------------------------------------------------
// keys: "exact_functionals" 0|1 (0 = reference bug-compatible potentials, default; 1 = potentials
//       that are the exact derivatives of the energies, i.e. libxc/PySCF numint)
//       "path" 0 auto | 1 generic (any alignment) | 2 TMA-fed DMMA kernels | 3 single-pass small-basis kernel (nao <= 48;
//       what auto picks there)
//       "deterministic" 0|1 (default 1: fixed-order reductions, bit-reproducible results)
//       "vxc_shape" 0|64|128|160 (tuning: output tile of the TMA V kernel -- 64 x 64, 128 x 128, 160 x 80;
//       0 = chosen from nao)
//       "vxc_vk" 0|8|16 (tuning: grid rows per ring stage of the 128 x 128 V kernel; 0 = 16 on dense
//       operands, 8 with zero skipping)
//       "small_streaming" 0|1 (small-basis kernel: 1 = always stream tile by tile; default 0 keeps whole 32-point
//       super-blocks resident in shared memory where two of them fit 16 KB per warp -- nao <= 8 for GGA, <= 32 for LDA)
//       "ao_shape" 0|1|8|16|17|32 (tuning: grid points per block of DFT_EvalAO -- 8 | 16 | 32 with 8 | 8 | 16 warps,
//       17 = 16 points with 16 warps; 0 = chosen from the basis size; 1 = the barrier-free direct kernel, lanes over
//       AOs throughout, measured slower)
//       "ao_vec_stores" 0|1 (DFT_EvalAO phase 2: 16-byte stores, a lane owns an aligned AO pair; default 0 -- measured
//       slower than 8-byte stores, whose warps already fill whole 256-byte runs)
//       "ao_input_order" 0|1 (DFT_EvalAO: 1 keeps the exponent-sharing shell groups in input order; default 0 sorts them
//       by reach and position so that the two half-warps of a warp decide alike on the cutoff)
//       "zero_skip" 0|1 (AO screening inside the contraction kernels: k-steps whose operand fragment is
//       exactly zero are skipped; results are unchanged; default 1)
//       "vxc_skip" -1|0|1 (the zero-skipping instance of the V kernel: -1 = adaptive, used while the density
//       kernel of the previous call skipped >= 10 % of its k-steps; default -1)
//       "vxc_skip_mode" 1|2|4 (which zero-skipping V instance on the 128 x 128 tile.  Per-warp votes on the M-side
//       fragments: 2 = built and voted on for a whole ring stage at once (default, C5 10.8 ms), 1 = k-step by k-step
//       (round 1, 11.1 ms); 4 = staged B: builder warps combine the planes once per CTA and all MMA warps skip the same
//       all-zero fragments (12.5 ms).  Diagnostic builds (-DDFT_V_EXPERIMENTS) also carry the measured-and-rejected
//       variants 3, 5, 6, 7: DESIGN.md 5.2e)
//       "vxc_rebalance" 0|1 (per-warp-vote instances: after every blocking call re-deal the 8-column M fragments to
//       the warps from that call's live counts -- heaviest with lightest; results are bit-identical; default 1)
//       "vxc_prefetch" n (tuning: L2 prefetch distance of the V kernel's producer in ring stages; default 0: measured
//       no gain)
//       "density_wide" 0|1 (density kernel: 1 = one consumer group of 8 warps on 128-point blocks with one seven-stage ring,
//       the Dsym chunk delivered once per 128 rows; 0 = two ping-pong groups of 4 warps on 64-point blocks)
//       "density_scatter" 0|1 (density kernel: visit the 64-point blocks in a scattered order -- golden-ratio stride --
//       instead of grid order, so that the SMs are not all in a sparse or a dense stretch of the grid together; default 0:
//       measured no gain)
//       "density_producers" 1|2 (tuning: TMA-issuing threads per consumer group of the density kernel, default 1:
//       a helper thread that brings the Dsym chunks and half of a piece's planes measured no gain)
//       "vxc_producers" 1..4 (tuning: TMA-issuing threads per CTA of the V kernel, default 2)
//       "vxc_scatter" 0|1 (zero-skipping V instances: scatter consecutive ring stages over the grid, default 1)
//       "dyn_sched" 0|1 (density kernel: hand the units of work out dynamically, default 1)
//       "density_unit" 0|1|2 (density kernel, unit of work: 1 = a 64-point block, 0 | 2 = one column tile of a
//       block, the default)
//       "stagger_min" n (density kernel: consumer group 1 starts half a tile period late when a CTA has more
//       than n blocks to do, default 8)
//       "wait_ns" n (tuning: producer threads sleep n ns between barrier polls, default 0)
//       "raw_convention" 0|1 (default 0: d_vxc is always the symmetric matrix S with 1/2 (S + S^T) = V_xc, which
//       is all the reference's driver uses (dft.py:212); 1: GGA leaves the reference's own raw, UNSYMMETRISED
//       B^T Phi with the 4 v_sigma factor (dft_solver.cu:616) for a caller that reads d_vxc directly.  LDA and
//       B3LYP raw outputs are already symmetric in the reference and identical in both settings)
//       ("debug_nodmma" exists only in -DDFT_DIAGNOSTICS builds, see the end of this header)
//       "tma_3d" 0|1 (tuning: 3-D tensor maps in the V kernel, one TMA load per plane and stage; default 1)
//       "l2_prefetch" 0|1 (tuning: short-range L2 prefetch in the density kernel, default 0: measured no gain)
//       "timing" 0|1 (record the per-kernel CUDA events behind DFT_GetStat "*_ms", default 0: at H2O size five event
//       records are a tenth of a call; bench.py and the tools switch it on where they read the times)
int DFT_SetOption(XCSolver* solver, const char* key, double value);
// keys: "density_ms", "vxc_ms", "reduce_ms", "total_ms" (CUDA-event times of the last
//       DFT_ComputeXC on the engine's stream), "launches" (kernels launched by the last call),
//       "ao_ms" (kernel time of the last DFT_EvalAO), "skip_fraction" (share of the density kernel's
//       k-steps that were exact zeros and skipped in the last call: the AO-screening statistic),
//       "vxc_skip_fraction" (staged-B V kernel: share of its (8-column fragment, k-step) units -- 2 DMMAs each --
//       that were skipped as exact zeros in the last call),
//       "path" (path actually taken), "workspace_bytes", "nranks" (ranks of the communicator, 1 without
//       DFT_CommInit), "plans_built" (TMA launch plans encoded so far: a steady SCF loop over the same arrays
//       builds exactly one), "dyn_units" (draws from the density kernel's dynamic work counter in the last call =
//       units of work + consumer groups when "dyn_sched" is on, 0 with the static deal), "density_units" /
//       "density_groups" (units of work and consumer groups of the density kernel's last launch).
double DFT_GetStat(XCSolver* solver, const char* key);

// ---- Coulomb and exact exchange in one pass over the ERI ------------
// d_J (nao,nao) = sum_kl (ij|kl) D[k,l], what DFT_ComputeCoulomb writes (dft_solver.cu:550-555), and
// d_K (nao,nao) = sum_jl (ij|kl) D[j,l], what the reference's driver computes for B3LYP with
// cupy.einsum('ijkl,jl->ik', eri, dm) (dft.py:218) in a second pass over the 8 nao^4-byte ERI.
// Every ERI element is loaded once; J relies on (ij|kl) = (kl|ij), which every ERI tensor has (the
// reference's column-major gemv reads the transposed matrix).  Enqueued on the engine stream (like DFT_ComputeCoulomb: no host
// sync); returns 0, or non-zero on bad arguments / CUDA failure.
int DFT_ComputeCoulombExchange(XCSolver* solver, int nao, unsigned long long d_eri_ptr,
                               unsigned long long d_dm_ptr, unsigned long long d_J_ptr,
                               unsigned long long d_K_ptr);

// ---- device-resident Fock assembly (host<->device chatter) ---------------
// The reference's driver downloads J, V_xc and K every iteration and assembles the Fock matrix on the
// host (dft.py:210-223, :230-236).  With these two calls only F travels (for the host eigh):
//   DFT_BuildFock:   d_F = Hcore + J + 1/2 (V_raw + V_raw^T) - 1/2 c_hf K     (d_K_ptr may be 0)
//   DFT_SCFEnergies: out3 = { sum D o Hcore, 1/2 sum D o J, -1/4 c_hf sum D o K }  (host array, blocking)
int DFT_BuildFock(XCSolver* solver, int nao, unsigned long long d_hcore_ptr, unsigned long long d_J_ptr,
                  unsigned long long d_vxc_ptr, unsigned long long d_K_ptr, double c_hf,
                  unsigned long long d_F_ptr);
int DFT_SCFEnergies(XCSolver* solver, int nao, unsigned long long d_dm_ptr, unsigned long long d_hcore_ptr,
                    unsigned long long d_J_ptr, unsigned long long d_K_ptr, double c_hf, double* out3);

// ---- convenience for callers that keep V_xc/E_xc on the device (no host sync) --------------
// Same as DFT_ComputeXC but writes E_xc to the device double at d_exc_ptr, does not block the
// host, and returns 0.  Work is enqueued on the engine stream; DFT_StreamSynchronize waits.
int DFT_ComputeXCAsync(XCSolver* solver, int ngrid, int nao,
                       unsigned long long d_dm_ptr, unsigned long long d_ao_ptr,
                       unsigned long long d_ao_grad_ptr, unsigned long long d_weights_ptr,
                       unsigned long long d_vxc_ptr, unsigned long long d_exc_ptr);
int DFT_StreamSynchronize(XCSolver* solver);
// Raw cudaStream_t of the engine (as an integer) so a harness can record CUDA events on it.
unsigned long long DFT_GetStream(XCSolver* solver);

// ---- roofline denominators measured in place (bench.py) -------------------------------------
// Register-resident FP64 tensor-core (mma.sync m8n8k4 f64 -> SASS DMMA) and FP64 FMA peak
// throughput in TFLOP/s on the current device; `iters` inner iterations per thread.
double DFT_MicrobenchDMMA(int iters);
double DFT_MicrobenchDFMA(int iters);
// FP64 tensor throughput (TFLOP/s) with exactly `warps_per_sm` (1..8) resident warps per SM: shows how many
// warps per SM sub-partition the DMMA pipe needs to saturate.
double DFT_MicrobenchDMMAWarps(int warps_per_sm, int iters);

const char* DFT_B200_Version(void);

#ifdef DFT_DIAGNOSTICS
// ---- diagnostic builds only (python -m quantum_compute_dft_b200.build --diag -> weights/dft_diag.so) ----------
// Not part of the product library: weights/dft.so neither exports DFT_DebugRead nor accepts the option
// "debug_nodmma" (TMA kernels run their whole pipeline but skip every DMMA -- RESULTS ARE WRONG; it measures the
// operand-delivery floor).  DFT_DebugRead copies an engine workspace ("coef", "epart", "dsym", "vpart", "rho",
// "scratch") to the host.
int DFT_DebugRead(XCSolver* solver, const char* what, void* dst, unsigned long long nbytes);
#endif
}
