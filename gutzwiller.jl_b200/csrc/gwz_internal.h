// gwz_internal.h -- host-side interface between the C ABI (gwz_api.cu) and the kernel launchers.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "gwz_device.cuh"
#include "gwz_generic.cuh"

namespace gwz {

constexpr int NUM_ACC = 13;        // doubles in the cross-device accumulator vector (GWZ_NUM_ACC)
constexpr int NUM_WALKER_ACC = 5;  // per-walker running sums kept in HBM: st, stE, stG, stGE, stEE (shifted)
constexpr int NUM_BACC = 38;       // blocking accumulators (GWZ_NUM_BACC): see BAccSlot
constexpr int BLOCKING_BLOCK = 128;
constexpr int WALKER_BLOCK = STEP_BLOCK;
constexpr int SWEEP_BLOCK = 128;

// accumulator vector layout (all-reduced with ncclSum)
enum AccSlot {
    ACC_WALKERS = 0,  // number of walkers
    ACC_VAL = 1,      // sum_w val_w                      (state.jl:137-149)
    ACC_VAL2 = 2,     // sum_w val_w^2                    (cross-walker error)
    ACC_GRAD = 3,     // sum_w grad_w
    ACC_GRAD2 = 4,    // sum_w grad_w^2
    ACC_T = 5,        // sum tau
    ACC_TE = 6,       // sum tau*E
    ACC_TG = 7,       // sum tau*G
    ACC_TGE = 8,      // sum tau*G*E
    ACC_TEE = 9,      // sum tau*E^2
    ACC_HOPS = 10,    // hop evaluations of this call
    ACC_STEPS = 11,   // walker-steps of this call
    ACC_VARW = 12     // sum_w of the per-walker time-weighted variance of E_loc (state.jl:293-298)
};

// accumulators of the per-walker blocking analyses (all-reduced with ncclSum next to the vector above)
enum BAccSlot {
    BACC_WALKERS = 0,  // walkers analysed
    BACC_MEAN = 1,     // sum_w b_w.mean                       (state.jl:150-164)
    BACC_ERR2 = 2,     // sum over the successful walkers of b_w.err^2
    BACC_EE2 = 3,      // ... of b_w.err_err^2                 (state.jl:279-283)
    BACC_FAIL = 4,     // walkers whose M test failed (k = -1)
    BACC_GRAD = 5,     // sum_w grad_w with val = b_w.mean      (state.jl:57-58)
    BACC_KHIST = 6     // 32 slots: number of walkers accepted at level k (k >= 31 counted in the last slot)
};

struct WalkerLaunch {
    DevModel dm;
    DevTables tb;
    int nw;    // 32-bit words of the extended string
    int kind;  // gwz_kernel_kind
    // walker state (SoA in HBM)
    uint64_t* addr;    // [W][n] unary bitstrings
    uint32_t* planes;  // [(NP)(NWM)][n] occupation bit-planes (gwz_planes.cuh), when the walkers live there
    double* acc;       // [NUM_WALKER_ACC][n]
    size_t n;
    uint64_t global_offset, seed, step0, steps;
    int accumulate;  // 0: burn-in (no estimator update, no reduction)
    int reset_acc;   // 1: clear the per-walker sums before the first step
    double shift_e, shift_g;  // (E_ref, G_ref): reference point of the shifted running sums
    // optional per-step series [steps][n] (recording run)
    double* rec_tau;
    double* rec_eloc;
    uint64_t* rec_addr;  // [steps][n][W]
    int32_t* rec_hop;    // [steps][n]: 2 z + right of the selected hop (walker_fast_kernel only)
    const double* u01s;  // [steps][n]: uniforms to use instead of the Philox stream (walker_fast_kernel only)
    // per-block partial accumulator vectors [grid][NUM_ACC]
    double* partials;
    int grid;
    int* err_flag;
};

// occupancy-derived persistent grid for the walker kernel of this (nw, kind, record)
cudaError_t walker_grid(int nw, int N, int kind, int kc, bool record, size_t n, int sm_count, int* grid);
cudaError_t launch_walkers(const WalkerLaunch& L, cudaStream_t s);
// out[ncomp] = sum over blocks of partials[block][ncomp], fixed order
// mirror (optional): pinned host copy of out, written by the kernel; err_dev -> err_host likewise
cudaError_t launch_reduce_partials(const double* partials, int nblocks, int ncomp, double* out, cudaStream_t s,
                                   double* mirror = nullptr, const int* err_dev = nullptr, int* err_host = nullptr);
cudaError_t launch_walker_init(uint64_t* addr, double* acc, size_t n, int W, const uint64_t* start_dev, cudaStream_t s);
cudaError_t launch_walker_estimates(const double* acc, size_t n, double shift_e, double* val_w, double* grad_w,
                                    cudaStream_t s);
// shift[0..1] = (E_loc, grad ratio) of walker 0's current address
cudaError_t launch_walker_ref(const DevModel& dm, const DevTables& tb, int nw, const uint64_t* addr, size_t n,
                              double* shift, cudaStream_t s);
// occupation bit-plane path (gwz_walker_planes.cu)
void planes_shape(int N, int M, int* nwm, int* np);
cudaError_t walker_planes_grid(int N, int M, size_t n, int sm_count, int* grid);
cudaError_t launch_walkers_planes(const WalkerLaunch& L, cudaStream_t s);
// incremental class-count path on the same plane format (gwz_walker_incr.cu)
bool walker_incr_supported(int N, int M);
cudaError_t walker_incr_grid(int N, int M, size_t n, int sm_count, int* grid);
cudaError_t launch_walkers_incr(const WalkerLaunch& L, cudaStream_t s);
// round-2 production kernel: integer hop counts per rate class, table-driven updates (gwz_walker_fast.cu)
bool walker_fast_supported(int N, int M);
cudaError_t walker_fast_grid(int N, int M, size_t n, int sm_count, int* grid);
cudaError_t launch_walkers_fast(const WalkerLaunch& L, cudaStream_t s);
cudaError_t launch_unary_to_planes(const uint64_t* addr, size_t n, const DevModel& dm, uint32_t* planes, cudaStream_t s);
// walkers first..first+count-1 -> unary SoA addr[W][count]
cudaError_t launch_planes_to_unary(const uint32_t* planes, size_t n, const DevModel& dm, size_t first, size_t count,
                                   uint64_t* addr, cudaStream_t s);
// AoS (n x W) <-> SoA ([W][n]) address layout conversion
cudaError_t launch_transpose_addr(const uint64_t* src, uint64_t* dst, size_t n, int W, bool to_soa, cudaStream_t s);

struct ProbeLaunch {
    DevModel dm;
    DevTables tb;
    int nw, kind;
    const uint64_t* addr;  // AoS n x W
    size_t n;
    const double* u01;  // may be null
    double *e_loc, *tau, *grad;
    int32_t* n_hops;
    double* rates;  // n x 2M or null
    int32_t* chosen;
    uint64_t* next_addr;  // AoS
    int* err_flag;
};
cudaError_t launch_probe(const ProbeLaunch& L, cudaStream_t s);
cudaError_t launch_ansatz(const DevModel& dm, double g, int nw, const uint64_t* addr, size_t n, double* val,
                          double* grad, cudaStream_t s);

struct SweepLaunch {
    DevModel dm;
    DevTables tb;
    int nw;
    const uint64_t* basis;  // SoA [W][dim] or null => ranked enumeration
    uint64_t dim;           // states handled by this sweep
    uint64_t rank_begin;    // first rank (ranked mode)
    const uint64_t* binom;  // [65][65] binomial table in HBM (ranked mode)
    int with_grad;
    int symmetric;  // ranked mode: evaluate one representative per translation orbit (N+M <= 64)
    double* partials;  // [grid][4]
    int grid;
    uint64_t chunk;  // consecutive ranks per thread (ranked mode)
    uint64_t extra_state;  // symmetric mode: one more state evaluated with weight 1 (the uniform state, see sweep_run)
    int has_extra;
    const uint32_t* weight;  // stored mode, optional: orbit sizes of a representative basis
    // optional fused reduction (sweep_kernel, small grids): the last CTA sums the partials into out (+ pinned mirror)
    unsigned* done;
    double* out;
    double* mirror;
    // optional per-state output (terms kernel): out[count][4] for states first..first+count-1
    double* terms_out;
    uint64_t* states_out;  // AoS count x W
    uint64_t first, count;
};
cudaError_t sweep_grid(int nw, int kc, int N, int M, bool ranked, uint64_t dim, int sm_count, int* grid, uint64_t* chunk);
cudaError_t launch_sweep(const SweepLaunch& L, cudaStream_t s);
cudaError_t launch_sweep_collect(int N, int M, int B, uint64_t dim, uint64_t rank_begin, int grid, uint64_t chunk,
                                 const uint64_t* binom, unsigned long long* counter, uint64_t* store, uint32_t* mult,
                                 cudaStream_t s);
cudaError_t sort_representatives(uint64_t* keys, uint32_t* vals, uint64_t n, cudaStream_t s);
cudaError_t launch_sweep_terms(const SweepLaunch& L, cudaStream_t s);

// ---- per-walker resample + blocking analysis of recorded series (gwz_blocking.cu) -------------
struct BlockingLaunch {
    const double* tau;   // [steps][stride]
    const double* eloc;  // [steps][stride]
    size_t stride, n;    // walkers per row of the series arrays; walkers analysed (first n columns)
    uint64_t steps, len; // series length; resample length (state.jl:183 `len`)
    const double* acc;   // [NUM_WALKER_ACC][acc_stride] running sums of the same steps, or null
    size_t acc_stride;
    double shift_e, shift_g;
    double* scratch;     // rows of blocking_scratch_stride(len) doubles: the resampled series pass through here once.
                         // One row per RESIDENT THREAD (grid x BLOCKING_BLOCK rows, reused from walker to walker: the
                         // round trip stays in L2), or one per walker when the caller wants the resampled series back
    bool scratch_per_walker;
    double* out;         // [5][out_stride]: mean, err, err_err, k, blocks per walker, or null
    size_t out_stride;
    double* partials;    // [grid][NUM_BACC]
    int grid;
};
cudaError_t blocking_grid(uint64_t len, size_t n, int sm_count, int* grid);
size_t blocking_scratch_stride(uint64_t len);
cudaError_t launch_series_blocking(const BlockingLaunch& L, cudaStream_t s);

// ---- DVec(result) histogram (gwz_hist.cu) ----------------------------------------------------
cudaError_t launch_hist_fill(uint64_t* tkeys, double* tvals, size_t slots, cudaStream_t s);
cudaError_t launch_hist_insert(const uint64_t* keys, const double* tau, size_t cells, uint64_t* tkeys, double* tvals,
                               uint64_t mask, cudaStream_t s);
cudaError_t launch_hist_compact(const uint64_t* tkeys, const double* tvals, size_t slots, uint64_t* out_keys,
                                double* out_vals, unsigned long long* count, uint64_t cap, cudaStream_t s);

// ---- ansatz-generic engine (gwz_generic.cu) ------------------------------------------------
struct GenProbeLaunch {
    GenModel gm;
    const uint64_t* addr;  // AoS n x W
    size_t n;
    const double* u01;
    double *val, *grad;                 // ansatz(addr, p), val_and_grad(ansatz, addr, p): n, n x P
    double *e_loc, *tau, *grad_ratio;   // kinetic_sample!: n, n, n x P
    int32_t* n_hops;
    double* rates;  // n x 2M, reference hop order
    int32_t* chosen;
    uint64_t* next_addr;
    int* err_flag;
};
struct GenWalkerLaunch {
    GenModel gm;
    uint8_t* occ;  // [n][Mpad]
    int Mpad;
    double* acc;  // [n][3 + 2P]
    size_t n;
    uint64_t global_offset, seed, step0, steps;
    int accumulate, reset_acc;
    const double* shift;  // [1 + P] reference point (E, G) of the shifted sums
    double* partials;     // [grid * GEN_WARPS][gen_vec_len(P)]
    double *rec_tau, *rec_eloc;
    uint64_t* rec_addr;
    int grid;
    int* err_flag;
};
struct GenSweepLaunch {
    GenModel gm;
    const uint64_t* basis;  // SoA [W][dim] or null => ranked
    uint64_t dim, rank_begin;
    const uint64_t* binom;
    uint64_t chunk;
    int with_grad;
    double* partials;  // [grid * GEN_WARPS][2 + 2P]
    double* terms_out;
    uint64_t* states_out;
    uint64_t first, count;
    int grid;
    const uint32_t* weight;  // stored mode, optional: orbit sizes of a representative basis
};
int gen_ppl(int P);  // parameter slots per lane (1, 2, 4, 8), 0 if P > 256
int gen_group(int M, int P);  // lanes per walker of the walker kernel (8, 16 or 32)
cudaError_t gen_walker_grid(const GenModel& gm, size_t n, int sm_count, int* grid, int* walkers_per_cta);
cudaError_t gen_sweep_grid(const GenModel& gm, size_t n, int sm_count, int* grid, int* states_per_cta);
cudaError_t launch_gen_unary_to_occ(const uint64_t* addr, size_t n, const GenModel& gm, int Mpad, uint8_t* occ,
                                    bool broadcast, cudaStream_t s);
cudaError_t launch_gen_occ_to_unary(const uint8_t* occ, size_t n, const GenModel& gm, int Mpad, uint64_t* addr,
                                    cudaStream_t s);
cudaError_t launch_gen_probe(const GenProbeLaunch& L, cudaStream_t s);
cudaError_t launch_gen_walkers(const GenWalkerLaunch& L, cudaStream_t s);
cudaError_t launch_gen_estimates(const double* acc, size_t n, int P, const double* shift, double* val_w, double* grad_w,
                                 cudaStream_t s);
cudaError_t launch_gen_sweep(const GenSweepLaunch& L, cudaStream_t s);
// CombinationAnsatz (gwz_generic_combo.cu): same launch descriptors, L.gm is ignored
cudaError_t combo_walker_grid(const GenCombo& cb, size_t n, int sm_count, int* grid);
cudaError_t combo_sweep_grid(const GenCombo& cb, size_t n, int sm_count, int* grid);
cudaError_t launch_combo_probe(const GenProbeLaunch& L, const GenCombo& cb, cudaStream_t s);
cudaError_t launch_combo_walkers(const GenWalkerLaunch& L, const GenCombo& cb, cudaStream_t s);
cudaError_t launch_combo_sweep(const GenSweepLaunch& L, const GenCombo& cb, cudaStream_t s);

}  // namespace gwz
