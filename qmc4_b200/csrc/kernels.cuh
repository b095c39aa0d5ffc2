// kernels.cuh — device-side state and the launch interface between capi.cu and the kernels.
#pragma once
#include "plan.h"
#include <cuda_runtime.h>

// Dynamic shared memory of a kernel / a <<<>>> launch.  The QMC_EMULATE branches exist only for the
// CPU thread emulation of tests/emu (test infrastructure; never compiled into libqmc4b200.so).
#ifdef QMC_EMULATE
#define QMC_DYNAMIC_SMEM(name) unsigned char *name = emu::tls.dyn_smem
#define QMC_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emu::launch_lambda(dim3(grid), dim3(block), (size_t)(smem), [&]() { kernel(__VA_ARGS__); })
#else
#define QMC_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define QMC_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif

namespace qmc {

// One walker buffer (HBM, SoA over walkers; each walker's dofs doubles are contiguous in the
// reference's own per-walker layout so a group streams them with coalesced loads).
struct WalkerBuf {
  double *R;        // [cap*dofs]
  double *F;        // [cap*dofs]  grad log psi
  double *logpsi;   // [cap]
  double *E;        // [cap]       local energy of the configuration
};

constexpr int MAX_RANKS = 8;
// per-step reduction vector: {sum E, population, one-hot populations[MAX_RANKS], one-hot free walker slots[MAX_RANKS]}.
// One all-reduce (sum) of it gives every rank the E_T inputs (branching.cpp:75-76) and what the migration planner needs
// (ptools.cpp:244-268 allgathers the populations): 144 bytes.
constexpr int NSUM = 2 + 2 * MAX_RANKS;

// Device-resident control block: everything the step sequence needs lives here so that
// qmc_step() enqueues kernels without ever reading back (population, E_T, counters).
struct Ctl {
  int n_cur;                   // walkers in the current population (input of the next propagate)
  int overflow;                // population exceeded capacity
  int invalid_init;            // a constraint was violated during initial evaluation
  int n_prev;                  // physical slots [0,n_prev) of the current buffer are referenced by the parent map
  unsigned long long step;     // global step counter (Philox counter word)
  double eshift;               // E_T in use  (branching.cpp:65-97)
  double lsum[NSUM];           // local  {sum E, population, one-hot population vector, one-hot free slots}
  double gsum[NSUM];           // global {sum E, population, populations per rank, free slots per rank} (after all-reduce)
  double blk_esum, blk_ecount; // energy estimator (scalarAccumulator, accumulators.h:21-47)
  double blk_moves, blk_acc;   // metropolis counters (metropolis.cpp:24-26)
  double blk_steps;
  double est_weight;           // walkers accumulated into the histogram estimators
};

struct TraceDev {              // optional per-walker trace (device pointers, may be null)
  double *xi, *Rprop, *Fprop, *logpsi, *lap, *V, *Eprop, *q, *u_m, *u_b;
  int *accepted, *valid, *n_desc, *parent;
};

// inter-CTA workspace of the branching scan (kernels_scan.cu); zero between launches (the kernel resets it)
struct ScanWorkspace {
  int *ticket, *done;          // arrival / completion counters
  int *total, *nacc;           // per CTA: descendants of its range, accepted moves
  unsigned *flag;              // per CTA: total published
  double *esum;                // per CTA: sum of the energies of its descendants
};

struct StepArgs {
  WalkerBuf in, out;
  int *parent;                 // [cap] source slot in `in` for each slot of `out` (rewritten by the scan)
  int *n_desc;                 // [cap]
  double *partials;            // [grid*groups*2] moves / accepted per group
  Ctl *ctl;
  const double *coef;          // spline coefficient pool
  double *hist;                // histogram replicas [grid][hist_stride] or null
  int accumulate;              // sampling step: estimators on
  int cap;
  TraceDev trace;
  // pipeline 2 (kernels_v2.cu)
  int *flags;                  // [cap] bit 0: move accepted (the scan counts the Metropolis statistics from these)
  int first, count;            // move kernel: walkers [first, first+count); count < 0: [0, ctl->n_cur).  first = -1 - d: ctl->n_cur - d
  int n_adjust;                // scan: population moved this step = ctl->n_cur + n_adjust (arrivals - departures of the migration)
  ScanWorkspace ws;
};

struct EvalArgs {
  const double *R;             // [nW*dofs]
  int nW;
  double *logpsi, *F, *lap, *V, *E;   // any may be null
  int *valid;
  const double *coef;
  double *hist;                // replicas or null
  Ctl *ctl;
  int flag_invalid;            // set ctl->invalid_init when a configuration is invalid
};

struct LaunchCfg {
  int grid, block, nwarps_per_group, groups_per_cta;
  size_t smem;
};

// kernel selectors
enum : int { KSEL_GG_STEP = 0, KSEL_GG_NONE = 1, KSEL_GAUSS_NONE = 2, KSEL_GENERIC = 3, KSEL_COUNT = 4 };

cudaError_t launch_eval(int ksel, bool hist, const LaunchCfg &cfg, cudaStream_t st, const Plan &plan, const EvalArgs &a);
cudaError_t launch_propagate(int ksel, bool hist, const LaunchCfg &cfg, cudaStream_t st, const Plan &plan,
                             const DmcParams &dmc, const StepArgs &a);
cudaError_t launch_scan(cudaStream_t st, const DmcParams &dmc, const StepArgs &a, int n_partials, int multi_rank, int init_mode);
int scan_max_blocks();
cudaError_t launch_finalize_shift(cudaStream_t st, const DmcParams &dmc, Ctl *ctl);
cudaError_t launch_init_identity(cudaStream_t st, int *parent, int *n_desc, int n, Ctl *ctl);
cudaError_t launch_reduce_hist(cudaStream_t st, const double *hist, int replicas, int stride, double *out);
cudaError_t launch_tabulate(cudaStream_t st, const RadialFn &J, const double *coef, double minx, double deltax, long n, double *out);
cudaError_t launch_pack(cudaStream_t st, const WalkerBuf &buf, const int *parent, int dofs, int first, int count, double *slab);
cudaError_t launch_unpack(cudaStream_t st, const WalkerBuf &buf, int dofs, int first, int count, const double *slab);
cudaError_t configure_kernels(int max_smem_optin);
int propagate_regs(int ksel, bool hist);
int eval_regs(int ksel, bool hist);
int propagate_occupancy(int ksel, bool hist, size_t smem);
double measure_fp64_peak_tflops(cudaStream_t st, float *ms_out);

// pipeline 2 (kernels_v2.cu): the production step sequence
cudaError_t configure_kernels_v2(int max_smem_optin);
int propagate_v2_occupancy(int ksel, size_t smem);
cudaError_t launch_propagate_v2(int ksel, const LaunchCfg &cfg, int grid, cudaStream_t st, const Plan &plan, const DmcParams &dmc,
                                const StepArgs &a, double *e_old);
cudaError_t launch_scan_v2(cudaStream_t st, const DmcParams &dmc, const StepArgs &a, const double *e_old, int multi_rank);
cudaError_t launch_pack_tail(cudaStream_t st, const WalkerBuf &buf, const int *parent, const Ctl *ctl, int dofs, int n_send, double *slab);
cudaError_t launch_unpack_append(cudaStream_t st, const WalkerBuf &buf, int *parent, const Ctl *ctl, int dofs, int n_send, int n_recv,
                                 const double *slab);
cudaError_t launch_pack_moved(cudaStream_t st, const WalkerBuf &out, const double *e_old, const int *flags, const Ctl *ctl, int dofs, int n_send,
                              double *slab);
cudaError_t launch_unpack_moved(cudaStream_t st, const WalkerBuf &out, double *e_old, int *flags, const Ctl *ctl, int dofs, int n_send, int n_recv,
                                const double *slab);
cudaError_t launch_population_checksum(cudaStream_t st, const WalkerBuf &buf, const int *parent, const Ctl *ctl, unsigned long long *out);
cudaError_t launch_estimators(const LaunchCfg &cfg, cudaStream_t st, const Plan &plan, const WalkerBuf &buf, const int *parent,
                              const Ctl *ctl, double *hist, int occ, int grid);
int estimator_occupancy(int occ, size_t smem);

// the hot kernel of the one-species / one-pair-pass shapes (kernels_pair.cu): move step with the tournament tile schedule
int pair_warps_per_walker(int n);
size_t pair_group_bytes(const Plan &plan, int nwarps);
size_t pair_cta_extra_bytes(int variant);
int pair_variant_threads(int variant);
void pair_exp_table(double *tab256);
cudaError_t configure_kernels_pair(int max_smem_optin);
int pair_move_occupancy(int vk, int occ, size_t smem);
int pair_move_regs(int vk, int occ);
cudaError_t launch_pair_move(int vk, int occ, const LaunchCfg &cfg, int grid, cudaStream_t st, const Plan &plan, const DmcParams &dmc,
                             const double *dev_tab, const StepArgs &a, double *e_old);

size_t pair_hist_smem(const Plan &plan);
int pair_hist_occupancy(const Plan &plan);
cudaError_t launch_pair_hist(int grid, cudaStream_t st, const Plan &plan, int e, const WalkerBuf &buf, const int *parent, const Ctl *ctl, double *hist);

// small systems (kernels_small.cu): several walkers per warp
int small_segment_width(int n);
int small_walkers_per_cta(int n);
int small_move_occupancy(int n);
int small_move_regs(int n);
cudaError_t launch_small_move(int n, int grid, cudaStream_t st, const Plan &plan, const DmcParams &dmc, const StepArgs &a, double *e_old);

// forward walking (kernels_fw.cu)
size_t fw_record_group_bytes(const Plan &plan, int max_sf_bins);
cudaError_t configure_kernels_fw(int max_smem_optin);
int fw_chunks();
cudaError_t launch_fw_record(cudaStream_t st, int grid, size_t group_bytes, const Plan &plan, const WalkerBuf &buf, const int *n_desc,
                             const Ctl *ctl, double *rows, int *nslots_out);
cudaError_t launch_fw_compose(cudaStream_t st, int grid, const int *parent, const int *comp_in, int *comp_out, const Ctl *ctl, int fresh,
                              int *level_slots);
cudaError_t launch_fw_chain(cudaStream_t st, const int *comp, const int *links, int depth, int cap, long e_last, int levels, int first, int count,
                            int *chain);
cudaError_t launch_fw_records(cudaStream_t st, int unpack, double *rows, int row_depth, int row_levels, int stride, double *rec, int width,
                              int *links, int depth, int levels, int cap, long e_last, const int *chain, int base, int *comp, int first,
                              int count, double *msg);
cudaError_t launch_fw_copy_links(cudaStream_t st, int grid, const int *comp, int *link, const Ctl *ctl);
cudaError_t launch_fw_accumulate(cudaStream_t st, int grid, const int *parent, const int *links, int depth, int cap, long event, int fw,
                                 const Ctl *ctl, int *cnt, const double *rows, const int *nslots, int stride, int offset, int bins,
                                 double *partial, double *sums, double *weight);
// superfluidFraction: centre-of-mass records per surviving slot + lagged winding products along the ancestor links
int sfl_chunk_slots();
cudaError_t launch_sfl_record(cudaStream_t st, int grid, const Plan &plan, const WalkerBuf &buf, const int *n_desc, const Ctl *ctl, double *rec);
cudaError_t launch_sfl_accumulate(cudaStream_t st, const double *rec, const int *links, const int *n_desc, const Ctl *ctl, int depth, int cap,
                                 int width, long event, int jmax, int offA, int offB, int lag_stride, double *partial, double *sums,
                                 double *weights);

} // namespace qmc
