// capi.cu — the C ABI declared in include/qmc4_b200.h: context, term-list compiler ("plan"),
// HBM buffers, step sequencing on one CUDA stream, NCCL plumbing.  No physics here except the
// derived-coefficient formulas of the reference's Jastrow constructors.
#include "../../include/qmc4_b200.h"
#include "kernels.cuh"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <time.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace qmc;

// ---- NCCL, resolved at run time (the library also has to load on boxes without NCCL) ----------
extern "C" {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId_t;
}
namespace {
struct NcclApi {
  void *h = nullptr;
  int (*GetUniqueId)(ncclUniqueId_t *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId_t, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*CommAbort)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool load(std::string &err)
  {
    if (h) return true;
    /* QMC4_NCCL_LIB: explicit path of the NCCL to bind (e.g. the torch-bundled one); tried first */
    const char *names[] = {getenv("QMC4_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names)
      {
        if (!n || !*n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
      }
    if (!h)
      {
        err = std::string("cannot load NCCL: ") + dlerror();
        return false;
      }
#define QMC_SYM(field, name)                                              \
  field = reinterpret_cast<decltype(field)>(dlsym(h, name));              \
  if (!field)                                                             \
    {                                                                     \
      err = std::string("NCCL symbol missing: ") + name;                  \
      return false;                                                       \
    }
    QMC_SYM(GetUniqueId, "ncclGetUniqueId")
    QMC_SYM(CommInitRank, "ncclCommInitRank")
    QMC_SYM(CommDestroy, "ncclCommDestroy")
    QMC_SYM(CommAbort, "ncclCommAbort")
    QMC_SYM(AllReduce, "ncclAllReduce")
    QMC_SYM(AllGather, "ncclAllGather")
    QMC_SYM(Send, "ncclSend")
    QMC_SYM(Recv, "ncclRecv")
    QMC_SYM(GroupStart, "ncclGroupStart")
    QMC_SYM(GroupEnd, "ncclGroupEnd")
    QMC_SYM(GetErrorString, "ncclGetErrorString")
#undef QMC_SYM
    return true;
  }
};
NcclApi g_nccl;
constexpr int NCCL_INT32 = 2, NCCL_FLOAT64 = 8, NCCL_SUM = 0;
std::string g_create_error;
} // namespace

// ---- context ----------------------------------------------------------------------------------
struct WfTerm { int kind; std::vector<int> sets; RadialFn J; double hs; };
struct PotTerm { int kind; std::vector<int> sets; std::vector<double> p; };
struct EstTerm { int kind; std::vector<int> sets; int bins; double minx, maxx; std::vector<double> extra; };

struct qmc_ctx {
  int device = 0;
  bool planning_only = false;      // device -1: term-list compiler only (host logic tests); every compute entry point refuses
  int sms = 0, max_smem = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  double lBox[3];
  std::vector<int> N;
  std::vector<WfTerm> wf;
  std::vector<PotTerm> pot;
  std::vector<EstTerm> est;
  std::vector<double> coefpool;
  bool planned = false;
  Plan plan;
  DmcParams dmc;
  int ksel = KSEL_GENERIC;
  bool has_hist = false;
  LaunchCfg cfg;
  int64_t cap = 0;
  WalkerBuf buf[2] = {};
  int cur = 0;
  int *parent = nullptr, *n_desc = nullptr;
  double *partials = nullptr;
  Ctl *ctl = nullptr;
  double *coef = nullptr;
  double *hist = nullptr, *hist_out = nullptr;
  int n_partials = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double device_ms = 0;
  long sample_phase = 0;
  // scratch for host-pointer entry points
  double *scr = nullptr;
  size_t scr_bytes = 0;
  int *scr_i = nullptr;
  size_t scr_i_bytes = 0;
  // multi rank
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  double *slab_send = nullptr, *slab_recv = nullptr;
  size_t slab_bytes = 0;
  long launches = 0;
  // per-launch timing of the dominant kernel (propagate): ring of event pairs resolved at sync points
  static constexpr int NEV = 256;
  cudaEvent_t pev0[NEV] = {}, pev1[NEV] = {};
  int pev_used = 0;
  double prop_ms = 0;
  long prop_launches_timed = 0;
  int time_kernels = 0;
  // pipeline 2 (default; kernels_v2.cu): branching decided in the scan kernel, estimators in their own kernel, the per-step
  // all-reduce and the walker migration on a second (high-priority) stream off the critical path.  QMC4_PIPELINE=1 selects
  // the fused pipeline of kernels.cu, which runs with genealogical records, forceEnergy and the traced step also carry.
  int pipeline = 2;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_scan = nullptr, ev_ar = nullptr, ev_mig = nullptr, ev_move = nullptr;
  bool ar_pending = false;
  int *flags = nullptr;             // [cap] accepted bit per walker (move kernel -> scan)
  double *h_sums = nullptr;         // pinned copy of Ctl::gsum after the last all-reduce: what the migration planner reads
  bool pops_pending = false;        // h_sums holds populations the planner has not looked at yet
  int mig_every = 1;                // plan a migration every mig_every steps ...
  double mig_band = 0.02;           // ... if max-min population exceeds this fraction of the mean (dead band), min 1 walker
  long mig_phase = 0, mig_events = 0, mig_out = 0, mig_in = 0;
  int slab_walkers = 0;             // walkers one migration slab holds (bounds a step's transfers)
  bool p2p_connected = false;       // every directed pair of ranks has exchanged a message (NCCL connects lazily: ~140 ms per pair)
  unsigned long long *d_checksum = nullptr;
  // the hot kernel of the one-species / one-pair-pass shapes (kernels_pair.cu) replaces the general move kernel when the plan allows
  ScanWorkspace scan_ws = {};
  int *scan_ws_i = nullptr;
  double *scan_ws_d = nullptr;
  // CUDA graph of two consecutive steps (the walker buffers ping-pong, so two steps restore the launch arguments): single rank,
  // every step a sampling step, no per-launch timers.  Replayed nSteps/2 times by qmc_step (QMC4_GRAPHS=0 disables).
#ifndef QMC_EMULATE
  cudaGraphExec_t step_graph = nullptr;
#endif
  bool use_graphs = true;
  long graph_launches_per_replay = 0;
  long graph_replays = 0;
  bool lean_hist = false;           // g(r) of the one-species shapes through pair_hist_kernel (kernels_pair.cu)
  int lean_hist_grid = 0;
  bool small = false;               // kernels_small.cu: several walkers per warp (one species, N <= 16, no triplet terms)
  int small_grid = 0;
  bool trip_cyclic = false;
  int trip_mod = 0;
  int *trip_tab = nullptr;
  bool lean = false;
  int lean_vk = V_NONE, lean_occ = 0;      /* lean_occ: launch-geometry variant of kernels_pair.cu */
  LaunchCfg lean_cfg = {};
  double *lean_tab = nullptr;       // 2^(j/256), j = 0..255 (device)
  // QMC4_TRACE=n: device/host timeline of the first n steps of every qmc_step call, printed to stderr (diagnostics)
  struct StepTrace { cudaEvent_t m0, m1, s1, e1, a1; double h_enter, h_wait0, h_wait1, h_exit; int mig; };
  std::vector<StepTrace> trace;
  int trace_n = 0, trace_used = 0;
  int est_occ = 4, est_grid = 0;      // estimator kernel of pipeline 2: resident CTAs per SM it is compiled for, its grid
  double *e_old = nullptr;
  // forward walking (kernels_fw.cu; correlationEstimator.cpp): ring of per-slot histogram rows + ancestor links
  struct FwEst { int target, fw, record_steps; double *sums = nullptr, *weight = nullptr; };
  std::vector<FwEst> fw_est;
  int max_sf_bins = 0;
  int fw_depth = 0;                 // histogram rows kept = max(fw)+1 events
  int fw_link_depth = 0;            // ancestor links kept = max(fw_depth, superfluid record steps) events
  long fw_event = 0;                // sampling events since the last upload
  double *fw_rows = nullptr;        // [depth][cap][hist_stride]
  int *fw_links = nullptr;          // [depth][cap]  slot at event e -> ancestor slot at event e-1
  int *fw_comp[2] = {nullptr, nullptr};
  int fw_comp_cur = 0;
  int *fw_cnt = nullptr, *fw_nslots = nullptr;
  double *fw_partial = nullptr;
  size_t fw_group_bytes = 0;
  // superfluidFraction (correlationEstimator.cpp:187-357): ring of per-slot centre-of-mass records on the same event clock
  struct SflEst { int setA, setB, bins; double *sums = nullptr, *weights = nullptr; };
  std::vector<SflEst> sfl_est;
  int sfl_record_steps = 0;         // centerOfMassStorer::recordSteps = "bins" of the first superfluidFraction measurement
  double *sfl_rec = nullptr;        // [fw_link_depth][cap][3*nSpecies]
  double *sfl_partial = nullptr;    // [chunks][record_steps]
  bool genealogy() const { return !fw_est.empty() || !sfl_est.empty(); }
  // migration of the records with their walker (qmc_rebalance): slots used per ring level, message scratch
  int *fw_level_slots = nullptr;    // [fw_link_depth] physical slots the records of that event occupy
  int *fw_chain = nullptr;
  double *fw_msg_send = nullptr, *fw_msg_recv = nullptr;
  size_t fw_chain_bytes = 0, fw_msg_bytes = 0;
  // forceEnergy (energy.cpp:37-61) for the methods without branching (main.cpp:318-335,395-408): a cold pass per sampling step
  bool force_energy = false;
  int64_t n_host = 0;               // population known to the host: constant while branching is off
  double *fe_R = nullptr, *fe_F = nullptr, *fe_V = nullptr, *fe_partial = nullptr, *fe_acc = nullptr;   // fe_acc = {sum, count}
};

namespace {

#define CK(call)                                                                              \
  do                                                                                          \
    {                                                                                         \
      cudaError_t e__ = (call);                                                               \
      if (e__ != cudaSuccess)                                                                 \
        {                                                                                     \
          c->err = std::string(#call) + ": " + cudaGetErrorString(e__);                       \
          return 1;                                                                           \
        }                                                                                     \
    }                                                                                         \
  while (0)
#define NK(call)                                                                              \
  do                                                                                          \
    {                                                                                         \
      int e__ = (call);                                                                       \
      if (e__ != 0)                                                                           \
        {                                                                                     \
          c->err = std::string(#call) + ": " + g_nccl.GetErrorString(e__);                    \
          return 1;                                                                           \
        }                                                                                     \
    }                                                                                         \
  while (0)

int fail(qmc_ctx *c, const std::string &m)
{
  c->err = m;
  return 1;
}

// Derived coefficients, following the reference constructors (file:line in each case).
int derive_radial(qmc_ctx *c, int kind, const double *p, int np, RadialFn &J)
{
  memset(&J, 0, sizeof J);
  J.kind = kind;
  auto need = [&](int n) { return np >= n; };
  switch (kind)
    {
    case QMC_J_GAUSSIAN: /* jastrowGaussian.h:12-18 */
      if (!need(1)) return fail(c, "gaussian: {alpha}");
      J.p[0] = p[0];
      return 0;
    case QMC_J_GAUSSGAUSS: /* jastrowGaussGauss.cpp:5-17: alpha = 1/(2 r0^2) */
      {
        if (!need(2)) return fail(c, "gaussiangaussian: {r0,u0}");
        const double alpha = 1. / (2 * p[0] * p[0]);
        J.p[0] = alpha;
        J.p[1] = p[1];
        const long double a = -(long double)alpha;
        J.p[2] = (double)(a * 1.442695040888963407359924681001892137L);
        const long double c_ = 0.693147180559945309417232121458176568L / a;
        double hi = (double)c_;
        uint64_t bits;
        memcpy(&bits, &hi, 8);
        bits &= 0xffffffff00000000ull;          /* 21 significant bits: n*hi exact for |n| < 2^31 */
        memcpy(&hi, &bits, 8);
        J.p[3] = hi;
        J.p[4] = (double)(c_ - (long double)hi);
        static const double T[11] = {1.0, 1.0, 0.5, 0.16666666666648303, 0.041666666666651364, 0.008333333353717156,
                                     0.0013888888905871347, 0.0001984120875699232, 2.4801536409064087e-05,
                                     2.7625102005388108e-06, 2.761379555451986e-07};
        long double ap = 1;
        for (int j = 1; j <= 10; ++j)
          {
            ap *= a;
            J.p[4 + j] = (double)((long double)T[j] * ap);
          }
        return 0;
      }
    case QMC_J_SQUAREWELL: /* jastrowSquareWell.cpp:5-44 ; JSON ctor: (V0,R0,Rm,alpha,aInverse,cut_off*2) */
      {
        if (!need(6)) return fail(c, "squareWell: {V0,R0,Rm,alpha,aInverse,cut_off}");
        const double V0 = p[0], R0 = p[1], Rm = p[2], alpha = p[3], aInv = p[4], lBox = p[5] * 2;
        const double K0 = std::sqrt(V0), half = lBox / 2.;
        const double A = std::sin(K0 * R0) / (1 - R0 * aInv);
        const double C = -A / (Rm * Rm * alpha * (std::exp(-alpha * (lBox - Rm)) - std::exp(-alpha * Rm)));
        const double B = A * (1 / Rm - aInv) - C * (std::exp(-alpha * Rm) + std::exp(-alpha * (lBox - Rm)));
        if (B + C * 2 * exp(-alpha * half) < 0) return fail(c, "Jastrow cannot be negative");
        J.p[0] = R0; J.p[1] = Rm; J.p[2] = half; J.p[3] = K0; J.p[4] = aInv; J.p[5] = alpha; J.p[6] = lBox;
        J.p[7] = B; J.p[8] = C; J.p[9] = std::log(B + C * 2 * std::exp(-alpha * half));
        return 0;
      }
    case QMC_J_SQUAREWELL_BOUND: /* jastrowSquareWell.cpp:46-95 ; JSON ctor: (V0,R0,Rm,cut_off*2,aInverse,eb,alpha) */
      {
        if (!need(7)) return fail(c, "squareWellBoundState: {V0,R0,Rm,cut_off,aInverse,eb,alpha}");
        const double V0 = p[0], R0 = p[1], Rm = p[2], lBox = p[3] * 2., eb = p[5], alpha = p[6];
        const double m = 0.5, K0 = sqrt(2 * m * (V0 - eb)), half = lBox / 2., ec = sqrt(2 * m * eb);
        const double A = exp(-ec * R0) / sin(K0 * R0);
        const double C = exp(-ec * Rm) * (1. / (Rm * Rm) + ec / Rm) / (alpha * (exp(-alpha * Rm) - exp(-alpha * (lBox - Rm))));
        const double B = -C * (exp(-alpha * Rm) + exp(-alpha * (lBox - Rm))) + exp(-ec * Rm) / Rm;
        if (B + C * 2 * exp(-alpha * half) < 0) return fail(c, "Jastrow cannot be negative");
        J.p[0] = R0; J.p[1] = Rm; J.p[2] = half; J.p[3] = K0; J.p[4] = log(A); J.p[5] = ec; J.p[6] = alpha; J.p[7] = lBox;
        J.p[8] = B; J.p[9] = C; J.p[10] = log(B + C * 2 * exp(-alpha * half));
        return 0;
      }
    case QMC_J_DELTA_PHONONS: /* jastrow_delta.cpp:5-39: parameters as given, k2 = pi/(2 cut_off) */
      if (!need(5)) return fail(c, "delta_phonons: {k,delta,beta,z,cut_off}");
      J.p[0] = p[0]; J.p[1] = p[1]; J.p[2] = p[2]; J.p[3] = p[3]; J.p[4] = p[4]; J.p[5] = M_PI / (2 * p[4]);
      return 0;
    case QMC_J_DIPOLAR_REP: /* jastrowDipolar.cpp:3-27: D = sqrt(j["D"]), lBox = 2 cut_off, k = pi/lBox, logC = log(C) */
      if (!need(5)) return fail(c, "dipolar_rep: {D,cut_off,matching_point,alpha,C}");
      J.p[0] = sqrt(p[0]); J.p[1] = (p[1] * 2) / 2.; J.p[2] = p[2]; J.p[3] = p[3]; J.p[4] = M_PI / (p[1] * 2); J.p[5] = log(p[4]);
      return 0;
    case QMC_J_DELTA_IN_TRAP: /* jastrow_delta.cpp:71-74 */
      if (!need(1)) return fail(c, "delta_in_trap: {a}");
      J.p[0] = p[0];
      return 0;
    case QMC_J_DELTA_BOUND_STATE_PHONONS: /* jastrow_delta.cpp:76-86: lBox = 2 cut_off, k2 = pi/lBox */
      if (!need(5)) return fail(c, "delta_bound_state_phonons: {xI,k,A,beta,cut_off}");
      J.p[0] = p[0]; J.p[1] = p[1]; J.p[2] = log(p[2]); J.p[3] = p[3]; J.p[4] = M_PI / (p[4] * 2);
      return 0;
    case QMC_J_HARDSPHEREGAUSS: /* jastrowHardSphereGauss.cpp:4-16 */
      {
        if (!need(3)) return fail(c, "hardSphereGauss: {a,Rm,D}");
        const double a = p[0], Rm = p[1], d = p[2];
        const double alpha = a / (Rm * (Rm - a) * 2 * (d - Rm));
        const double C = (1 - a / Rm) * exp(alpha * (Rm - d) * (Rm - d));
        J.p[0] = Rm; J.p[1] = a; J.p[2] = d; J.p[3] = alpha; J.p[4] = log(C);
        return 0;
      }
    case QMC_J_HARDSPHERE: /* jastrowHardSphereGauss.cpp:18-32 */
      {
        if (!need(3)) return fail(c, "hardSphere: {a,k,cut_off}");
        J.p[0] = p[0]; J.p[1] = p[2]; J.p[2] = p[2] / sin(p[1] * (p[2] - p[0])); J.p[3] = p[1];
        return 0;
      }
    case QMC_J_POSCHTELLER: /* jastrowPoschTeller.cpp:4-24 ; longDistanceConstant has no log (sic) */
      {
        if (!need(5)) return fail(c, "poschTeller: {R0,C,alpha,Rm,cut_off}");
        const double R0 = p[0], C = p[1], alpha = p[2], Rm = p[3], lBox = p[4] * 2.;
        J.p[0] = Rm; J.p[1] = lBox / 2.; J.p[2] = 1 / R0; J.p[3] = alpha; J.p[4] = lBox; J.p[5] = C;
        J.p[6] = 2 * C * exp(-alpha * lBox / 2.);
        return 0;
      }
    case QMC_J_BSPLINE: /* jastrowSpline.cpp:20-78 */
      {
        if (!need(3 + 4)) return fail(c, "bSpline: {stepSize,valueRight,derivativeRight,coefficients...}");
        const int nc = np - 3;
        J.ncoef = nc;
        J.coef_off = (int)c->coefpool.size();
        for (int i = 0; i < nc; ++i) c->coefpool.push_back(p[3 + i]);
        J.p[0] = 1. / p[0];
        J.p[1] = (nc - 3) * p[0];
        J.p[2] = p[1];
        J.p[3] = p[2];
        return 0;
      }
    default:
      return fail(c, "unknown jastrow kind " + std::to_string(kind));
    }
}

// largest double t with sqrt(t) <= R0 (sqrt correctly rounded and monotone), so that
//   r2 <= t   <=>   sqrt(r2) <= R0   bit for bit (potentialFunctors.h:9,26 compare the rounded norm)
double step_threshold(double R0)
{
  if (!(R0 > 0)) return -1.0;
  double t = R0 * R0;
  while (std::sqrt(t) > R0) t = std::nextafter(t, 0.0);
  while (std::sqrt(std::nextafter(t, INFINITY)) <= R0) t = std::nextafter(t, INFINITY);
  return t;
}

bool same_pair(int a0, int a1, int b0, int b1) { return (a0 == b0 && a1 == b1) || (a0 == b1 && a1 == b0); }

int build_plan(qmc_ctx *c)
{
  Plan &P = c->plan;
  memset(&P, 0, sizeof P);
  const int S = (int)c->N.size();
  P.nSpecies = S;
  int off = 0;
  for (int s = 0; s < S; ++s)
    {
      P.N[s] = c->N[s];
      P.off[s] = off;
      off += 3 * c->N[s];
      P.Ntot += c->N[s];
    }
  P.dofs = off;
  for (int d = 0; d < 3; ++d)
    {
      P.L[d] = c->lBox[d];
      P.Linv[d] = 1. / c->lBox[d];       /* geometry.h:39 */
    }
  auto set_ok = [&](int s) { return s >= 0 && s < S; };
  // estimators first (pair passes refer to them)
  int hoff = 0;
  for (size_t e = 0; e < c->est.size(); ++e)
    {
      if (P.nest >= MAX_EST) return fail(c, "too many estimators");
      const EstTerm &E = c->est[e];
      Estimator &D = P.est[P.nest++];
      D.kind = E.kind;
      D.setA = E.sets[0];
      D.setB = E.sets.size() > 1 ? E.sets[1] : E.sets[0];
      D.setC = E.sets.size() > 2 ? E.sets[2] : -1;
      if (!set_ok(D.setA) || !set_ok(D.setB)) return fail(c, "estimator: invalid set");
      const bool three_sets = E.kind == QMC_EST_TRIMERCORRELATION || E.kind == QMC_EST_TRIMERCORRELATION_DIS || E.kind == QMC_EST_TRIMERHYPERRADIUS;
      if (three_sets)
        {
          if (E.sets.size() < 3 || !set_ok(D.setC)) return fail(c, "three-set observable needs sets={A,B,C}");
          /* asserts of trimerCorrelation.cpp:11,46-47,202 and hyperRadius.cpp:6-7,23 */
          if (D.setA == D.setB || D.setA == D.setC || D.setB == D.setC) return fail(c, "three-set observable: the sets must be different");
          if (E.kind != QMC_EST_TRIMERCORRELATION_DIS && P.N[D.setC] != 1)
            return fail(c, "trimerCorrelation / trimerhyperRadius: set C must hold exactly one particle");
        }
      if ((E.kind == QMC_EST_TRIMERHYPERRADIUS || E.kind == QMC_EST_CENTEROFMASSSQUARED) && E.bins != 1)
        return fail(c, "scalar observable: bins must be 1");
      D.bins = E.bins;
      D.offset = hoff;
      hoff += E.bins;
      D.minx = E.minx;
      D.maxx = E.maxx;
      const double dx = (E.maxx - E.minx) / E.bins;       /* accumulators.h:108 */
      D.dxinv = 1. / dx;
      if (E.kind == QMC_EST_PAIRCORRELATION)
        { /* pairCorrelation.cpp:41-59 */
          double nf = std::pow(c->lBox[0], 3) / (dx * 4 * M_PI * P.N[D.setA] * P.N[D.setB]);
          D.norm = (D.setA == D.setB) ? 2 * nf : nf;
          D.fx_scale = std::ldexp(1.0, 36) * E.maxx * E.maxx;     /* 1/r^2 at r = maxx -> 2^36 */
        }
      else if (E.kind == QMC_EST_TRIMERCORRELATION_UNDIS)
        { /* trimerCorrelation.cpp:128-143 */
          const double NA = P.N[D.setA];
          D.norm = 1 / (dx * 4 * M_PI * NA * NA * NA);
        }
      else if (E.kind == QMC_EST_TRIMERCORRELATION)
        { /* trimerCorrelation.cpp:58-78: NC = getN(setB) (sic) */
          const double NA = P.N[D.setA], NB = P.N[D.setB];
          D.norm = 1 / (dx * 4 * M_PI * NA * NB * NB);
        }
      else if (E.kind == QMC_EST_TRIMERCORRELATION_DIS)
        { /* trimerCorrelation.cpp:209-227 */
          const double NA = P.N[D.setA], NB = P.N[D.setB], NC = P.N[D.setC];
          D.norm = 1 / (dx * 4 * M_PI * NA * NB * NC);
        }
      else if (E.kind == QMC_EST_TRIMERHYPERRADIUS || E.kind == QMC_EST_CENTEROFMASSSQUARED)
        {
          D.norm = 1;
          D.minx = 0;
          D.maxx = 1;
          D.dxinv = 1;
        }
      else if (E.kind == QMC_EST_STRUCTUREFACTOR)
        { /* structureFactor.cpp:8-40 */
          if (E.extra.size() < 3) return fail(c, "structureFactor: extra={maxq,lBox,channel}");
          const double qMax = E.extra[0], lb = E.extra[1];
          double dq = int(qMax / (2 * M_PI * E.bins / lb)) * 2 * M_PI / lb;
          if (dq == 0) dq = 2 * M_PI / lb;
          D.q0 = 2 * M_PI / lb;
          D.dq = dq;
          D.channel = (int)E.extra[2];
          D.norm = 1;
        }
      else return fail(c, "unknown estimator kind");
    }
  P.hist_stride = hoff;
  // wavefunction terms
  for (const WfTerm &w : c->wf)
    {
      for (int s : w.sets)
        if (!set_ok(s)) return fail(c, "wavefunction term: invalid set");
      if (w.kind == QMC_WF_JASTROW1B)
        {
          if (P.n1b >= MAX_ONEBODY) return fail(c, "too many one-body terms");
          OneBodyTerm &T = P.t1[P.n1b++];
          T.set = w.sets[0];
          T.J = w.J;
        }
      else if (w.kind == QMC_WF_JASTROW2B_UNDIS || w.kind == QMC_WF_JASTROW2B_DIS)
        {
          if (w.sets.size() < 2) return fail(c, "two-body jastrow needs sets={A,B}");
          if (w.kind == QMC_WF_JASTROW2B_UNDIS && w.sets[0] != w.sets[1])
            return fail(c, "setA != setB in undistinguishable two body jastrow");   /* jastrowWavefunctionTwoBody.h:100 */
          if (w.kind == QMC_WF_JASTROW2B_DIS && w.sets[0] == w.sets[1])
            return fail(c, "setA == setB in distinguishable two body jastrow");     /* jastrowWavefunctionTwoBody.h:11 */
          if (P.n2b >= MAX_PAIRPASS) return fail(c, "too many pair passes");
          PairPass &T = P.p2[P.n2b++];
          T.setA = w.sets[0];
          T.setB = w.sets[1];
          T.vkind = V_NONE;
          T.hist = -1;
          T.J = w.J;
        }
      else if (w.kind == QMC_WF_JASTROW3B_UNDIS)
        {
          if (P.n3b >= MAX_TRIPLETPASS) return fail(c, "too many triplet passes");
          TripletPass &T = P.p3[P.n3b++];
          T.set = w.sets[0];
          T.setB = T.setC = -1;
          T.hist = -1;
          T.J = w.J;
          if (w.hs >= 0)
            {
              T.has_constraint = 1;
              T.hs_R0 = w.hs;
              T.hs_thr2 = step_threshold(w.hs);
            }
        }
      else if (w.kind == QMC_WF_JASTROW3B_DIS)
        {
          if (w.sets.size() < 3) return fail(c, "jastrow3bDis needs sets={A,B,C}");
          if (w.sets[0] == w.sets[1] || w.sets[1] == w.sets[2] || w.sets[0] == w.sets[2])
            return fail(c, "Sets in three body wavefunction should be distinguishible");   /* jastrowWavefunctionThreeBody.h:173-176 */
          if (P.n3b >= MAX_TRIPLETPASS) return fail(c, "too many triplet passes");
          TripletPass &T = P.p3[P.n3b++];
          T.set = w.sets[0];
          T.setB = w.sets[1];
          T.setC = w.sets[2];
          T.hist = -1;
          T.J = w.J;
          if (w.hs >= 0)
            {
              T.has_constraint = 1;
              T.hs_R0 = w.hs;
            }
        }
      else return fail(c, "unknown wavefunction kind");
    }
  // potentials
  for (const PotTerm &v : c->pot)
    {
      for (int s : v.sets)
        if (!set_ok(s)) return fail(c, "potential term: invalid set");
      if (v.kind == QMC_POT_HARMONIC1B)
        {
          OneBodyTerm *T = nullptr;
          for (int t = 0; t < P.n1b; ++t)
            if (P.t1[t].set == v.sets[0] && !P.t1[t].has_pot) T = &P.t1[t];
          if (!T)
            {
              if (P.n1b >= MAX_ONEBODY) return fail(c, "too many one-body terms");
              T = &P.t1[P.n1b++];
              T->set = v.sets[0];
              T->J.kind = J_NONE;
            }
          T->has_pot = 1;
          T->half_omega2 = 0.5 * v.p[0] * v.p[0];      /* potential.cpp:29 */
        }
      else if (v.kind == QMC_POT_THREEBODY_UNDIS_SQUAREWELL || v.kind == QMC_POT_THREEBODY_UNDIS_BARRIER)
        {
          TripletPass *T = nullptr;
          for (int t = 0; t < P.n3b; ++t)
            if (P.p3[t].setB < 0 && P.p3[t].set == v.sets[0] && !P.p3[t].has_pot) T = &P.p3[t];
          if (!T)
            {
              if (P.n3b >= MAX_TRIPLETPASS) return fail(c, "too many triplet passes");
              T = &P.p3[P.n3b++];
              T->set = v.sets[0];
              T->setB = T->setC = -1;
              T->hist = -1;
              T->J.kind = J_NONE;
            }
          T->has_pot = 1;
          T->v_R0 = v.p[0];
          T->v_thr2 = step_threshold(v.p[0]);
          T->v_val = (v.kind == QMC_POT_THREEBODY_UNDIS_SQUAREWELL) ? -v.p[1] : v.p[1];   /* potentialFunctors.h:9,26 */
        }
      else if (v.kind == QMC_POT_THREEBODY_DIS_SQUAREWELL || v.kind == QMC_POT_THREEBODY_DIS_BARRIER)
        {
          if (v.sets.size() < 3) return fail(c, "potentialThreeBodyDis needs sets={A,B,C}");
          if (v.sets[0] == v.sets[1] || v.sets[1] == v.sets[2] || v.sets[0] == v.sets[2])
            return fail(c, "Sets in three body wavefunction should be distinguishible");   /* potentialThreeBody.h:68-71 (sic) */
          TripletPass *T = nullptr;
          for (int t = 0; t < P.n3b; ++t)   /* same ordered triple: the three pair tables and the loop order coincide */
            if (P.p3[t].set == v.sets[0] && P.p3[t].setB == v.sets[1] && P.p3[t].setC == v.sets[2] && !P.p3[t].has_pot && !T) T = &P.p3[t];
          if (!T)
            {
              if (P.n3b >= MAX_TRIPLETPASS) return fail(c, "too many triplet passes");
              T = &P.p3[P.n3b++];
              T->set = v.sets[0];
              T->setB = v.sets[1];
              T->setC = v.sets[2];
              T->hist = -1;
              T->J.kind = J_NONE;
            }
          T->has_pot = 1;
          T->v_R0 = v.p[0];
          T->v_val = (v.kind == QMC_POT_THREEBODY_DIS_SQUAREWELL) ? -v.p[1] : v.p[1];
        }
      else
        {
          if (v.sets.size() < 2) return fail(c, "two-body potential needs sets={A,B}");
          int vk = V_NONE;
          double thr = 0, val = 0;
          switch (v.kind)
            {
            case QMC_POT_SQUAREWELL2B: /* params {V0,R0}: -count(d<=R0)*V0, potential.cpp:57-65 */
              vk = V_STEP; thr = step_threshold(v.p[1]); val = -v.p[0]; break;
            case QMC_POT_TWOBODY_SQUAREWELL: /* params {R0,V0} */
              vk = V_STEP; thr = step_threshold(v.p[0]); val = -v.p[1]; break;
            case QMC_POT_TWOBODY_BARRIER:
              vk = V_STEP; thr = step_threshold(v.p[0]); val = v.p[1]; break;
            case QMC_POT_DIPOLAR2B:
              vk = V_DIPOLAR; val = v.p[0]; break;
            case QMC_POT_POSCHTELLER2B:
              vk = V_POSCHTELLER; val = 1 / v.p[0]; break;      /* alpha = 1/R0, potential.cpp:121 */
            default: return fail(c, "unknown potential kind");
            }
          PairPass *T = nullptr;
          for (int t = 0; t < P.n2b; ++t)
            if (same_pair(P.p2[t].setA, P.p2[t].setB, v.sets[0], v.sets[1]) && P.p2[t].vkind == V_NONE && !T) T = &P.p2[t];
          if (!T)
            {
              if (P.n2b >= MAX_PAIRPASS) return fail(c, "too many pair passes");
              T = &P.p2[P.n2b++];
              T->setA = v.sets[0];
              T->setB = v.sets[1];
              T->hist = -1;
              T->J.kind = J_NONE;
            }
          T->vkind = vk;
          T->v_thr2 = thr;
          T->v_val = val;
        }
    }
  // estimators attach to passes
  c->has_hist = P.nest > 0;
  int max_sf_bins = 0;
  for (int e = 0; e < P.nest; ++e)
    {
      Estimator &D = P.est[e];
      if (D.kind == QMC_EST_PAIRCORRELATION)
        {
          PairPass *T = nullptr;
          for (int t = 0; t < P.n2b; ++t)
            if (same_pair(P.p2[t].setA, P.p2[t].setB, D.setA, D.setB) && P.p2[t].hist < 0 && !T) T = &P.p2[t];
          if (!T)
            {
              if (P.n2b >= MAX_PAIRPASS) return fail(c, "too many pair passes");
              T = &P.p2[P.n2b++];
              T->setA = D.setA;
              T->setB = D.setB;
              T->vkind = V_NONE;
              T->J.kind = J_NONE;
            }
          T->hist = e;
        }
      else if (D.kind == QMC_EST_TRIMERCORRELATION_UNDIS)
        {
          TripletPass *T = nullptr;
          for (int t = 0; t < P.n3b; ++t)
            if (P.p3[t].setB < 0 && P.p3[t].set == D.setA && P.p3[t].hist < 0 && !T) T = &P.p3[t];
          if (!T)
            {
              if (P.n3b >= MAX_TRIPLETPASS) return fail(c, "too many triplet passes");
              T = &P.p3[P.n3b++];
              T->set = D.setA;
              T->setB = T->setC = -1;
              T->J.kind = J_NONE;
            }
          T->hist = e;
        }
      else if (D.kind == QMC_EST_STRUCTUREFACTOR) max_sf_bins = std::max(max_sf_bins, D.bins);
    }
  // put a pass with a specialised kernel first
  auto spec = [&](const PairPass &T) -> int {
    if (T.J.kind == J_GAUSSGAUSS && T.vkind == V_STEP) return KSEL_GG_STEP;
    if (T.J.kind == J_GAUSSGAUSS && T.vkind == V_NONE) return KSEL_GG_NONE;
    if (T.J.kind == J_GAUSSIAN && T.vkind == V_NONE) return KSEL_GAUSS_NONE;
    return (int)KSEL_GENERIC;
  };
  c->ksel = KSEL_GENERIC;
  for (int t = 0; t < P.n2b; ++t)
    if (spec(P.p2[t]) != KSEL_GENERIC)
      {
        std::swap(P.p2[0], P.p2[t]);
        c->ksel = spec(P.p2[0]);
        break;
      }
  if (const char *e = getenv("QMC4_FORCE_GENERIC"))
    if (atoi(e)) c->ksel = KSEL_GENERIC;
  // work decomposition: warps per walker group from the tile count of the largest pair pass
  int maxN = 1;
  for (int s = 0; s < S; ++s) maxN = std::max(maxN, P.N[s]);
  const int nb = (maxN + 31) / 32, tiles = nb * (nb + 1) / 2;
  int wpw = tiles <= 3 ? 1 : tiles <= 10 ? 2 : tiles <= 36 ? 4 : 8;
  if (P.n3b > 0) wpw = 8;
  /* cyclic triplet schedule (every triplet once): all undistinguishable triplet passes on sets of one size M, 3 <= M <= 32 */
  int trip_mod = 0;
  bool cyclic = P.n3b > 0;
  for (int t = 0; t < P.n3b; ++t)
    {
      const TripletPass &T = P.p3[t];
      if (T.setB >= 0) cyclic = false;
      else if (trip_mod == 0) trip_mod = P.N[T.set];
      else if (trip_mod != P.N[T.set]) cyclic = false;
    }
  if (trip_mod < 3 || trip_mod > 32) cyclic = false;
  for (int e = 0; e < P.nest; ++e)
    if (P.est[e].kind == QMC_EST_TRIMERCORRELATION || P.est[e].kind == QMC_EST_TRIMERCORRELATION_DIS || P.est[e].kind == QMC_EST_TRIMERHYPERRADIUS)
      cyclic = false;
  if (const char *e = getenv("QMC4_TRIPLET_CYCLIC"))
    if (atoi(e) == 0) cyclic = false;
  c->trip_cyclic = cyclic;
  c->trip_mod = cyclic ? trip_mod : 0;
  if (cyclic) wpw = 2;
  if (const char *e = getenv("QMC4_WPW"))
    {
      int v = atoi(e);
      if (v == 1 || v == 2 || v == 4 || v == 8) wpw = v;
    }
  int tab = 0;
  for (int t = 0; t < P.n3b; ++t)
    {
      const TripletPass &T = P.p3[t];
      const int n = P.N[T.set];
      if (T.setB < 0) tab = std::max(tab, 2 * (n * (n - 1) / 2));
      else tab = std::max(tab, 2 * (n * P.N[T.setB] + P.N[T.setB] * P.N[T.setC] + n * P.N[T.setC]));
    }
  for (int e = 0; e < P.nest; ++e)     /* three-set observables build their own r^2 tables (no weights) */
    if (P.est[e].kind == QMC_EST_TRIMERCORRELATION || P.est[e].kind == QMC_EST_TRIMERCORRELATION_DIS || P.est[e].kind == QMC_EST_TRIMERHYPERRADIUS)
      {
        const int a = P.N[P.est[e].setA], b = P.N[P.est[e].setB], cc = P.N[P.est[e].setC];
        tab = std::max(tab, a * b + b * cc + a * cc);
      }
  if (c->trip_cyclic) tab = std::max(tab, 1024 * (1 + wpw));      /* r2[32][32] + one weight table per warp */
  P.needs_pair_table = tab;
  P.aux_doubles = max_sf_bins * 2 * wpw;
  c->max_sf_bins = max_sf_bins;
  size_t group_doubles = (size_t)P.dofs * (1 + wpw) + (wpw * 4 + 8) + tab + P.aux_doubles;
  group_doubles = (group_doubles + 1) & ~(size_t)1;
  c->cfg.block = 256;
  c->cfg.nwarps_per_group = wpw;
  c->cfg.groups_per_cta = 8 / wpw;
  c->cfg.smem = group_doubles * 8 * c->cfg.groups_per_cta + (size_t)P.hist_stride * 8;   /* + per-CTA fixed-point histogram */
  if ((int)c->cfg.smem > c->max_smem)
    return fail(c, "configuration needs " + std::to_string(c->cfg.smem) + " B of shared memory per CTA (limit " +
                       std::to_string(c->max_smem) + "): too many particles for this build");
  return 0;
}

int ensure_scratch(qmc_ctx *c, size_t bytes, size_t ibytes)
{
  if (bytes > c->scr_bytes)
    {
      if (c->scr) cudaFree(c->scr);
      c->scr = nullptr;
      CK(cudaMalloc(&c->scr, bytes));
      c->scr_bytes = bytes;
    }
  if (ibytes > c->scr_i_bytes)
    {
      if (c->scr_i) cudaFree(c->scr_i);
      c->scr_i = nullptr;
      CK(cudaMalloc(&c->scr_i, ibytes));
      c->scr_i_bytes = ibytes;
    }
  return 0;
}

// SMs the compute kernels may fill.  With a communicator two SMs are left to the communication stream: a persistent move
// kernel that owns every CTA slot makes the all-reduce, the pack/unpack kernels and NCCL's point-to-point kernels wait for
// its END, and a collective that found a slot first displaces a move CTA whose whole share of walkers then starts late
// (measured at 8 GPUs: move kernel +9 %, step +7 %).  QMC4_COMM_SMS overrides.
int compute_sms(const qmc_ctx *c)
{
  int reserve = c->nranks > 1 ? 2 : 0;
  if (const char *e = getenv("QMC4_COMM_SMS")) reserve = std::max(0, std::min(16, atoi(e)));
  return std::max(1, c->sms - reserve);
}

int grid_for(qmc_ctx *c, int64_t nW)
{
  // persistent-style grid: a multiple of the SM count (resident CTAs per SM bounded by shared memory)
  // resident CTAs per SM from the occupancy calculator (registers, shared memory, threads) for the
  // heavier of the two propagate variants, so that the grid is exactly one resident wave
  int per_sm = std::min(propagate_occupancy(c->ksel, false, c->cfg.smem), propagate_occupancy(c->ksel, c->has_hist, c->cfg.smem));
  if (c->pipeline == 2) per_sm = propagate_v2_occupancy(c->ksel, c->cfg.smem);
  per_sm = std::max(1, std::min(per_sm, 2048 / c->cfg.block));
  if (const char *e = getenv("QMC4_CTAS_PER_SM"))
    {
      int v = atoi(e);
      if (v >= 1 && v <= 8) per_sm = v;
    }
  int64_t want = (nW + c->cfg.groups_per_cta - 1) / c->cfg.groups_per_cta;
  int64_t g = std::min<int64_t>(want, (int64_t)compute_sms(c) * per_sm);
  return (int)std::max<int64_t>(g, 1);
}

int allreduce_sums(qmc_ctx *c)
{
  if (c->nranks > 1)
    {
      /* the next propagate kernel derives E_T from gsum itself: one 80-byte collective per step, no extra launch */
      NK(g_nccl.AllReduce(c->ctl->lsum, c->ctl->gsum, NSUM, NCCL_FLOAT64, NCCL_SUM, c->comm, c->stream));
    }
  return 0;
}

int resolve_kernel_timers(qmc_ctx *c)
{ // call after a stream synchronisation
  for (int i = 0; i < c->pev_used; ++i)
    {
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, c->pev0[i], c->pev1[i]));
      c->prop_ms += ms;
      c->prop_launches_timed += 1;
    }
  c->pev_used = 0;
  return 0;
}

StepArgs make_step_args(qmc_ctx *c, int accumulate)
{
  StepArgs a;
  memset(&a, 0, sizeof a);
  a.in = c->buf[c->cur];
  a.out = c->buf[c->cur ^ 1];
  a.parent = c->parent;
  a.n_desc = c->n_desc;
  a.partials = c->partials;
  a.ctl = c->ctl;
  a.coef = c->coef;
  a.hist = (accumulate && c->has_hist) ? c->hist : nullptr;
  a.accumulate = accumulate;
  a.cap = (int)c->cap;
  a.ws = c->scan_ws;
  return a;
}

// wait (on the host) for everything either stream still has in flight
int sync_all(qmc_ctx *c)
{
  if (c->comm_stream) CK(cudaStreamSynchronize(c->comm_stream));
  CK(cudaStreamSynchronize(c->stream));
  c->ar_pending = false;
  return 0;
}

// Forward walking bookkeeping after the scan of a step (kernels_fw.cu).  c->parent is the parent map of the
// NEXT step (post-branching walker -> slot of the buffer just written), c->n_desc the descendants per slot.
int fw_after_step(qmc_ctx *c, int sample)
{
  const int cap = (int)c->cap, D = c->fw_depth, DL = c->fw_link_depth, stride = c->plan.hist_stride;
  const int g = std::max(1, std::min(c->sms * 4, (cap + 255) / 256));
  if (!sample)
    { /* between sampling events the ancestor map is composed with every step's parent map */
      CK(launch_fw_compose(c->stream, g, c->parent, c->fw_comp[c->fw_comp_cur], c->fw_comp[c->fw_comp_cur ^ 1], c->ctl, 0, nullptr));
      c->fw_comp_cur ^= 1;
      c->launches += 1;
      return 0;
    }
  const long e = c->fw_event;
  if (e > 0)
    {
      CK(launch_fw_copy_links(c->stream, g, c->fw_comp[c->fw_comp_cur], c->fw_links + (size_t)(e % DL) * cap, c->ctl));
      c->launches += 1;
    }
  if (!c->sfl_est.empty())
    { /* centerOfMassStorer::store, then superfluidFractionEstimator::accumulate (estimatorCollection.cpp:37-49) */
      const int width = 3 * c->plan.nSpecies, R = c->sfl_record_steps, jmax = (int)(e % R);
      CK(launch_sfl_record(c->stream, g, c->plan, c->buf[c->cur], c->n_desc, c->ctl, c->sfl_rec + (size_t)(e % DL) * cap * width));
      c->launches += 1;
      for (auto &f : c->sfl_est)
        {
          CK(launch_sfl_accumulate(c->stream, c->sfl_rec, c->fw_links, c->n_desc, c->ctl, DL, cap, width, e, jmax, 3 * f.setA, 3 * f.setB,
                                   R, c->sfl_partial, f.sums, f.weights));
          c->launches += 2;
        }
    }
  const int slot = c->fw_est.empty() ? 0 : (int)(e % D);
  if (!c->fw_est.empty())
    {
      CK(launch_fw_record(c->stream, std::max(1, std::min(cap, c->sms * 2)), c->fw_group_bytes, c->plan, c->buf[c->cur], c->n_desc, c->ctl,
                          c->fw_rows + (size_t)slot * cap * stride, c->fw_nslots + slot));
      c->launches += 1;
    }
  for (auto &f : c->fw_est)
    {
      if (e + 1 < f.record_steps) continue;     /* the walkers' rings are still filling (correlationEstimator.cpp:143) */
      const int tslot = (int)((e - f.fw) % D);
      const Estimator &E = c->plan.est[f.target];
      CK(launch_fw_accumulate(c->stream, g, c->parent, c->fw_links, DL, cap, e, f.fw, c->ctl, c->fw_cnt,
                              c->fw_rows + (size_t)tslot * cap * stride, c->fw_nslots + tslot, stride, E.offset, E.bins, c->fw_partial,
                              f.sums, f.weight));
      c->launches += 3;
    }
  CK(launch_fw_compose(c->stream, g, c->parent, nullptr, c->fw_comp[c->fw_comp_cur], c->ctl, 1, c->fw_level_slots + (e % DL)));
  c->launches += 1;
  c->fw_event = e + 1;
  return 0;
}

// The migration plan of one step, identical on every rank (all inputs come from the all-reduced vector in h_sums):
// alias-method plan (ptools.cpp:93-161) inside a dead band, clamped to the receivers' free slots and to the slab size.
void plan_migration(qmc_ctx *c, std::vector<int> &src, std::vector<int> &amt, int &n_send, int &n_recv)
{
  const int nr = c->nranks;
  std::vector<int> pops(nr), fre(nr), sent(nr, 0);
  src.assign(nr, 0);
  amt.assign(nr, 0);
  n_send = n_recv = 0;
  long total = 0;
  int mx = 0, mn = 1 << 30;
  for (int r = 0; r < nr; ++r)
    {
      pops[r] = (int)c->h_sums[2 + r];
      fre[r] = (int)c->h_sums[2 + MAX_RANKS + r];
      total += pops[r];
      mx = std::max(mx, pops[r]);
      mn = std::min(mn, pops[r]);
    }
  const double band = std::max(1.0, c->mig_band * (double)total / nr);
  if ((double)(mx - mn) <= band) return;
  qmc_plan_rebalance(pops.data(), nr, src.data(), amt.data());
  for (int r = 0; r < nr; ++r)
    if (amt[r] > 0)
      {
        const int a = std::min(std::min(amt[r], fre[r]), c->slab_walkers - sent[src[r]]);
        amt[r] = std::max(a, 0);
        sent[src[r]] += amt[r];
      }
  n_send = sent[c->rank];
  n_recv = amt[c->rank];
}

double host_now_ms()
{
  timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec * 1e3 + t.tv_nsec * 1e-6;
}

int one_step_v2(qmc_ctx *c, int accumulate, const TraceDev *tr)
{ /* move(k) || { all-reduce(k-1), migration(k-1) }  ->  [move(k) on the arrivals]  ->  scan(k)  ->  { all-reduce(k) } || estimators(k) || move(k+1) */
  qmc_ctx::StepTrace *T = (c->trace_used < c->trace_n) ? &c->trace[c->trace_used++] : nullptr;
  if (T)
    {
      T->h_enter = host_now_ms();
      T->h_wait0 = T->h_wait1 = 0;
      T->mig = 0;
      CK(cudaEventRecord(T->m0, c->stream));
    }
  StepArgs a = make_step_args(c, accumulate);
  a.hist = nullptr;
  a.flags = c->flags;
  a.first = 0;
  a.count = -1;
  a.n_adjust = 0;
  if (tr) a.trace = *tr;
  const bool timed = c->time_kernels && c->pev_used < qmc_ctx::NEV;
  if (timed) CK(cudaEventRecord(c->pev0[c->pev_used], c->stream));
  if (c->lean) CK(launch_pair_move(c->lean_vk, c->lean_occ, c->lean_cfg, c->lean_cfg.grid, c->stream, c->plan, c->dmc, c->lean_tab, a, c->e_old));
  else if (c->small) CK(launch_small_move(c->plan.N[0], c->small_grid, c->stream, c->plan, c->dmc, a, c->e_old));
  else CK(launch_propagate_v2(c->ksel, c->cfg, c->cfg.grid, c->stream, c->plan, c->dmc, a, c->e_old));
  if (timed) CK(cudaEventRecord(c->pev1[c->pev_used++], c->stream));
  if (T) CK(cudaEventRecord(T->m1, c->stream));
  c->launches += 1;
  if (c->nranks > 1 && c->dmc.branching_on && c->pops_pending && ++c->mig_phase % c->mig_every == 0)
    { /* one step late, like dmcDriver::isend (dmcDriver.cpp:112,150-157): while the move kernel runs, the host reads the
         populations the previous step's all-reduce delivered and posts the transfers on the communication stream */
      if (T) T->h_wait0 = host_now_ms();
      if (c->ar_pending) CK(cudaEventSynchronize(c->ev_ar));
      if (T) T->h_wait1 = host_now_ms();
      c->pops_pending = false;
        {
          std::vector<int> src, amt;
          int n_send = 0, n_recv = 0;
          plan_migration(c, src, amt, n_send, n_recv);
          if (n_send > 0 || n_recv > 0)
            {
              /* the tail walkers travel after their move of this step and before its branching: the sender has advanced them,
                 the receiver appends the moved state to its output population (nothing is moved twice, no extra launch) */
              const size_t dofs = c->plan.dofs, wsz = 2 * dofs + 4;
              CK(cudaEventRecord(c->ev_move, c->stream));
              CK(cudaStreamWaitEvent(c->comm_stream, c->ev_move, 0));
              CK(launch_pack_moved(c->comm_stream, a.out, c->e_old, c->flags, c->ctl, (int)dofs, n_send, c->slab_send));
              NK(g_nccl.GroupStart());
              size_t off = 0;
              for (int r = 0; r < c->nranks; ++r)
                if (amt[r] > 0 && src[r] == c->rank)
                  {
                    NK(g_nccl.Send(c->slab_send + off, (size_t)amt[r] * wsz, NCCL_FLOAT64, r, c->comm, c->comm_stream));
                    off += (size_t)amt[r] * wsz;
                  }
              if (n_recv > 0) NK(g_nccl.Recv(c->slab_recv, (size_t)n_recv * wsz, NCCL_FLOAT64, src[c->rank], c->comm, c->comm_stream));
              NK(g_nccl.GroupEnd());
              CK(launch_unpack_moved(c->comm_stream, a.out, c->e_old, c->flags, c->ctl, (int)dofs, n_send, n_recv, c->slab_recv));
              CK(cudaEventRecord(c->ev_mig, c->comm_stream));
              CK(cudaStreamWaitEvent(c->stream, c->ev_mig, 0));
              a.n_adjust = n_recv - n_send;
              c->launches += (n_send > 0) + (n_recv > 0);
              c->mig_events += 1;
              if (T) T->mig = n_send + n_recv;
              c->mig_out += n_send;
              c->mig_in += n_recv;
            }
        }
    }
  if (c->ar_pending) CK(cudaStreamWaitEvent(c->stream, c->ev_ar, 0));        /* scan needs E_T of the previous population */
  CK(launch_scan_v2(c->stream, c->dmc, a, c->e_old, c->nranks > 1));
  if (T) CK(cudaEventRecord(T->s1, c->stream));
  c->launches += 1;
  c->cur ^= 1;
  if (c->nranks > 1)
    {
      CK(cudaEventRecord(c->ev_scan, c->stream));
      CK(cudaStreamWaitEvent(c->comm_stream, c->ev_scan, 0));
      NK(g_nccl.AllReduce(c->ctl->lsum, c->ctl->gsum, NSUM, NCCL_FLOAT64, NCCL_SUM, c->comm, c->comm_stream));
      CK(cudaMemcpyAsync(c->h_sums, c->ctl->gsum, sizeof(double) * NSUM, cudaMemcpyDeviceToHost, c->comm_stream));
      CK(cudaEventRecord(c->ev_ar, c->comm_stream));
      if (T) CK(cudaEventRecord(T->a1, c->comm_stream));
      c->ar_pending = true;
      c->pops_pending = true;
    }
  else if (T) CK(cudaEventRecord(T->a1, c->stream));
  if (accumulate && c->has_hist)
    {
      if (c->lean_hist) CK(launch_pair_hist(c->lean_hist_grid, c->stream, c->plan, 0, c->buf[c->cur], c->parent, c->ctl, c->hist));
      else CK(launch_estimators(c->cfg, c->stream, c->plan, c->buf[c->cur], c->parent, c->ctl, c->hist, c->est_occ, c->est_grid));
      c->launches += 1;
    }
  if (T)
    {
      CK(cudaEventRecord(T->e1, c->stream));
      T->h_exit = host_now_ms();
    }
  return 0;
}

void dump_trace(qmc_ctx *c, double h0)
{
  if (c->trace_used == 0) return;
  if (c->comm_stream) cudaStreamSynchronize(c->comm_stream);
  fprintf(stderr, "[qmc4 trace rank %d] step: device ms since call start {move0 move1 scan1 est1 ar1} | host ms {enter wait0 wait1 exit} mig\n", c->rank);
  for (int i = 0; i < c->trace_used; ++i)
    {
      qmc_ctx::StepTrace &t = c->trace[i];
      float m0 = 0, m1 = 0, s1 = 0, e1 = 0, a1 = 0;
      cudaEventElapsedTime(&m0, c->ev0, t.m0);
      cudaEventElapsedTime(&m1, c->ev0, t.m1);
      cudaEventElapsedTime(&s1, c->ev0, t.s1);
      cudaEventElapsedTime(&e1, c->ev0, t.e1);
      cudaEventElapsedTime(&a1, c->ev0, t.a1);
      fprintf(stderr, "[qmc4 trace rank %d] %3d: %8.3f %8.3f %8.3f %8.3f %8.3f | %8.3f %8.3f %8.3f %8.3f  %d\n", c->rank, i, m0, m1, s1, e1, a1,
              t.h_enter - h0, t.h_wait0 ? t.h_wait0 - h0 : 0.0, t.h_wait1 ? t.h_wait1 - h0 : 0.0, t.h_exit - h0, t.mig);
    }
  c->trace_used = 0;
}

int force_energy_pass(qmc_ctx *c);

int one_step(qmc_ctx *c, int accumulate, const TraceDev *tr)
{
  if (c->pipeline == 2) return one_step_v2(c, accumulate, tr);
  StepArgs a = make_step_args(c, accumulate);
  if (tr) a.trace = *tr;
  const bool hist = accumulate && c->has_hist;
  const bool timed = c->time_kernels && c->pev_used < qmc_ctx::NEV;
  if (timed) CK(cudaEventRecord(c->pev0[c->pev_used], c->stream));
  CK(launch_propagate(c->ksel, hist, c->cfg, c->stream, c->plan, c->dmc, a));
  if (timed) CK(cudaEventRecord(c->pev1[c->pev_used++], c->stream));
  CK(launch_scan(c->stream, c->dmc, a, c->n_partials, c->nranks > 1, 0));
  c->launches += 2;
  c->cur ^= 1;
  if (allreduce_sums(c)) return 1;
  if (accumulate && c->force_energy && force_energy_pass(c)) return 1;
  return c->genealogy() ? fw_after_step(c, accumulate) : 0;
}

} // namespace

__global__ void gather_kernel(WalkerBuf buf, const int *parent, int dofs, int count, double *R, double *E)
{
  const int w = blockIdx.x;
  if (w >= count) return;
  const int p = parent[w];
  for (int k = threadIdx.x; k < dofs; k += blockDim.x) R[(size_t)w * dofs + k] = buf.R[(size_t)p * dofs + k];
  if (threadIdx.x == 0 && E) E[w] = buf.E[p];
}

// forceEnergy::operator() (energy.cpp:37-61): 0.5 * sum |grad log psi|^2 + V per walker; per-CTA running sums in a fixed order
constexpr int FE_THREADS = 128, FE_BLOCKS = 64;
__global__ void __launch_bounds__(FE_THREADS) force_energy_kernel(const double *F, const double *V, int dofs, int n, double *partial)
{
  __shared__ double sh[FE_THREADS];
  double total = 0.0;
  for (int w = blockIdx.x; w < n; w += gridDim.x)
    {
      double s = 0.0;
      for (int k = threadIdx.x; k < dofs; k += FE_THREADS)
        {
          const double f = F[(size_t)w * dofs + k];
          s += f * f;
        }
      sh[threadIdx.x] = s;
      __syncthreads();
      for (int h = FE_THREADS / 2; h > 0; h >>= 1)
        {
          if ((int)threadIdx.x < h) sh[threadIdx.x] += sh[threadIdx.x + h];
          __syncthreads();
        }
      if (threadIdx.x == 0) total += 0.5 * sh[0] + V[w];
      __syncthreads();
    }
  if (threadIdx.x == 0) partial[blockIdx.x] = total;
}
__global__ void force_energy_finish_kernel(const double *partial, int nb, int n, double *acc)
{
  if (threadIdx.x) return;
  double s = 0.0;
  for (int b = 0; b < nb; ++b) s += partial[b];
  acc[0] += s;
  acc[1] += (double)n;
}

extern "C" {
static int eval_common(qmc_ctx *c, const double *dR, int64_t nW, double *dlogpsi, double *dgrad, double *dlap,
                       double *dV, double *dE, int *dvalid, bool hist, int flag_invalid);
}
namespace {

// the observable re-evaluates the derivatives of the walker it is given (psi.evaluateDerivatives, energy.cpp:42): gather the
// population through the parent map, one evaluation launch, one reduction.  Never on the DMC path (branching on).
int force_energy_pass(qmc_ctx *c)
{
  const int n = (int)c->n_host, dofs = c->plan.dofs;
  if (n < 1) return 0;
  QMC_LAUNCH(gather_kernel, n, 128, 0, c->stream, c->buf[c->cur], c->parent, dofs, n, c->fe_R, (double *)nullptr);
  CK(cudaGetLastError());
  if (eval_common(c, c->fe_R, n, nullptr, c->fe_F, nullptr, c->fe_V, nullptr, nullptr, false, 0)) return 1;
  const int nb = std::min(n, FE_BLOCKS);
  QMC_LAUNCH(force_energy_kernel, nb, FE_THREADS, 0, c->stream, c->fe_F, c->fe_V, dofs, n, c->fe_partial);
  QMC_LAUNCH(force_energy_finish_kernel, 1, 32, 0, c->stream, c->fe_partial, nb, n, c->fe_acc);
  CK(cudaGetLastError());
  c->launches += 3;
  return 0;
}
} // namespace

// ================================================================================================
extern "C" {

int qmc_abi_version(void) { return QMC4_B200_ABI_VERSION; }

const char *qmc_last_error(const qmc_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int qmc_create(qmc_ctx **out, int device, const double lBox[3], int nSpecies, const int *N)
{
  if (!out) return 1;
  *out = nullptr;
  if (nSpecies < 1 || nSpecies > MAX_SPECIES)
    {
      g_create_error = "nSpecies must be in 1.." + std::to_string(MAX_SPECIES);
      return 1;
    }
  if (device == -1)
    { /* planning-only context: builds and exposes the kernel plan without touching CUDA (B200 limits assumed) */
      qmc_ctx *p = new qmc_ctx;
      p->device = -1;
      p->planning_only = true;
      p->sms = 148;
      p->max_smem = 232448;
      for (int d = 0; d < 3; ++d) p->lBox[d] = lBox[d];
      p->N.assign(N, N + nSpecies);
      memset(&p->dmc, 0, sizeof p->dmc);
      p->dmc.nranks = 1;
      *out = p;
      return 0;
    }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    {
      g_create_error = std::string("qmc4_b200 needs a CUDA device (no CPU fallback): ") + cudaGetErrorString(e);
      return 1;
    }
  if (device < 0 || device >= ndev)
    {
      g_create_error = "invalid device index";
      return 1;
    }
  qmc_ctx *c = new qmc_ctx;
  c->device = device;
  for (int d = 0; d < 3; ++d) c->lBox[d] = lBox[d];
  c->N.assign(N, N + nSpecies);
  auto bail = [&](const std::string &m) {
    g_create_error = m;
    delete c;
    return 1;
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(cudaGetErrorString(e));
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail(cudaGetErrorString(e));
  if (prop.major < 10) return bail(std::string("qmc4_b200 is built for sm_100a only; device is ") + prop.name);
  c->sms = prop.multiProcessorCount;
  c->max_smem = (int)prop.sharedMemPerBlockOptin;
  if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(cudaGetErrorString(e));
  if ((e = configure_kernels(c->max_smem)) != cudaSuccess) return bail(std::string("kernel image: ") + cudaGetErrorString(e));
  cudaEventCreate(&c->ev0);
  cudaEventCreate(&c->ev1);
  for (int i = 0; i < qmc_ctx::NEV; ++i)
    {
      cudaEventCreate(&c->pev0[i]);
      cudaEventCreate(&c->pev1[i]);
    }
  if (const char *pv = getenv("QMC4_PIPELINE"))
    if (atoi(pv) == 1) c->pipeline = 1;
  if ((e = configure_kernels_v2(c->max_smem)) != cudaSuccess) return bail(std::string("kernel image (v2): ") + cudaGetErrorString(e));
  if ((e = configure_kernels_pair(c->max_smem)) != cudaSuccess) return bail(std::string("kernel image (pair): ") + cudaGetErrorString(e));
  {
    /* the communication stream outranks the compute stream: the 144-byte all-reduce of step k is enqueued before the
       estimator / move kernels that follow the scan, and its CTA is placed first */
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if ((e = cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, hi)) != cudaSuccess) return bail(cudaGetErrorString(e));
    cudaEventCreateWithFlags(&c->ev_scan, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_ar, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_mig, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_move, cudaEventDisableTiming);
    if ((e = cudaHostAlloc((void **)&c->h_sums, sizeof(double) * NSUM, cudaHostAllocDefault)) != cudaSuccess) return bail(cudaGetErrorString(e));
    memset(c->h_sums, 0, sizeof(double) * NSUM);
  }
  if (const char *v = getenv("QMC4_TRACE"))
    {
      c->trace_n = std::max(0, std::min(512, atoi(v)));
      c->trace.resize(c->trace_n);
      for (auto &t : c->trace)
        {
          cudaEventCreate(&t.m0);
          cudaEventCreate(&t.m1);
          cudaEventCreate(&t.s1);
          cudaEventCreate(&t.e1);
          cudaEventCreate(&t.a1);
        }
    }
  if (const char *v = getenv("QMC4_GRAPHS"))
    c->use_graphs = atoi(v) != 0;
  if (const char *v = getenv("QMC4_MIGRATE_EVERY"))
    c->mig_every = std::max(1, atoi(v));
  if (const char *v = getenv("QMC4_MIGRATE_BAND"))
    c->mig_band = std::max(0.0, atof(v));
  if ((e = cudaMalloc(&c->ctl, sizeof(Ctl))) != cudaSuccess) return bail(cudaGetErrorString(e));
  cudaMemset(c->ctl, 0, sizeof(Ctl));
  memset(&c->dmc, 0, sizeof c->dmc);
  c->dmc.nranks = 1;
  *out = c;
  return 0;
}

void qmc_destroy(qmc_ctx *c)
{
  if (!c) return;
  if (c->planning_only)
    {
      delete c;
      return;
    }
  cudaSetDevice(c->device);
  if (c->comm_stream) cudaStreamSynchronize(c->comm_stream);
  if (c->stream) cudaStreamSynchronize(c->stream);
  cudaFree(c->e_old);
  cudaFree(c->flags);
  cudaFree(c->d_checksum);
  cudaFree(c->lean_tab);
#ifndef QMC_EMULATE
  if (c->step_graph) cudaGraphExecDestroy(c->step_graph);
#endif
  cudaFree(c->trip_tab);
  cudaFree(c->scan_ws_i);
  cudaFree(c->scan_ws_d);
  if (c->h_sums) cudaFreeHost(c->h_sums);
  if (c->ev_scan) cudaEventDestroy(c->ev_scan);
  if (c->ev_ar) cudaEventDestroy(c->ev_ar);
  if (c->ev_mig) cudaEventDestroy(c->ev_mig);
  if (c->ev_move) cudaEventDestroy(c->ev_move);
  if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
  for (int b = 0; b < 2; ++b)
    {
      cudaFree(c->buf[b].R);
      cudaFree(c->buf[b].F);
      cudaFree(c->buf[b].logpsi);
      cudaFree(c->buf[b].E);
    }
  cudaFree(c->parent);
  cudaFree(c->n_desc);
  cudaFree(c->partials);
  cudaFree(c->ctl);
  cudaFree(c->coef);
  cudaFree(c->hist);
  cudaFree(c->hist_out);
  cudaFree(c->scr);
  cudaFree(c->scr_i);
  cudaFree(c->slab_send);
  cudaFree(c->slab_recv);
  cudaFree(c->fw_rows);
  cudaFree(c->fw_links);
  cudaFree(c->fw_comp[0]);
  cudaFree(c->fw_comp[1]);
  cudaFree(c->fw_cnt);
  cudaFree(c->fw_nslots);
  cudaFree(c->fw_partial);
  for (auto &f : c->fw_est)
    {
      cudaFree(f.sums);
      cudaFree(f.weight);
    }
  cudaFree(c->fw_level_slots);
  cudaFree(c->fw_chain);
  cudaFree(c->fw_msg_send);
  cudaFree(c->fw_msg_recv);
  cudaFree(c->fe_R);
  cudaFree(c->fe_F);
  cudaFree(c->fe_V);
  cudaFree(c->fe_partial);
  cudaFree(c->fe_acc);
  cudaFree(c->sfl_rec);
  cudaFree(c->sfl_partial);
  for (auto &f : c->sfl_est)
    {
      cudaFree(f.sums);
      cudaFree(f.weights);
    }
  if (c->comm) g_nccl.CommDestroy(c->comm);
  for (int i = 0; i < qmc_ctx::NEV; ++i)
    {
      if (c->pev0[i]) cudaEventDestroy(c->pev0[i]);
      if (c->pev1[i]) cudaEventDestroy(c->pev1[i]);
    }
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int qmc_add_wavefunction_term(qmc_ctx *c, int wf_kind, const int *sets, int nsets, int jastrow_kind,
                              const double *params, int nparams, double hardSphereRadius)
{
  if (!c) return 1;
  if (c->planned) return fail(c, "terms must be added before qmc_set_dmc");
  if (nsets < 1) return fail(c, "sets missing");
  WfTerm t;
  t.kind = wf_kind;
  t.sets.assign(sets, sets + nsets);
  t.hs = hardSphereRadius;
  if (derive_radial(c, jastrow_kind, params, nparams, t.J)) return 1;
  c->wf.push_back(t);
  return 0;
}

int qmc_add_potential_term(qmc_ctx *c, int pot_kind, const int *sets, int nsets, const double *params, int nparams)
{
  if (!c) return 1;
  if (c->planned) return fail(c, "terms must be added before qmc_set_dmc");
  if (nsets < 1) return fail(c, "sets missing");
  const int need = (pot_kind == QMC_POT_HARMONIC1B || pot_kind == QMC_POT_DIPOLAR2B || pot_kind == QMC_POT_POSCHTELLER2B) ? 1 : 2;
  if (nparams < need) return fail(c, "potential: missing parameters");
  PotTerm t;
  t.kind = pot_kind;
  t.sets.assign(sets, sets + nsets);
  t.p.assign(params, params + nparams);
  c->pot.push_back(t);
  return 0;
}

int qmc_add_estimator(qmc_ctx *c, int est_kind, const int *sets, int nsets, int bins, double minx, double maxx,
                      const double *extra, int nextra, int *est_index)
{
  if (!c) return 1;
  if (c->planned) return fail(c, "estimators must be added before qmc_set_dmc");
  if (nsets < 1 || bins < 1) return fail(c, "estimator: sets/bins missing");
  EstTerm t;
  t.kind = est_kind;
  t.sets.assign(sets, sets + nsets);
  t.bins = bins;
  t.minx = minx;
  t.maxx = maxx;
  if (extra && nextra > 0) t.extra.assign(extra, extra + nextra);
  if (est_kind == QMC_EST_STRUCTUREFACTOR)
    {
      t.minx = 0;
      t.maxx = bins;
    }
  if (est_index) *est_index = (int)c->est.size();
  c->est.push_back(t);
  return 0;
}

int qmc_add_forward_walking(qmc_ctx *c, int est_index, int record_steps, const int *fw_steps, int n_fw, int *first_index)
{
  if (!c) return 1;
  if (c->planned) return fail(c, "forward walking must be declared before qmc_set_dmc");
  if (est_index < 0 || est_index >= (int)c->est.size()) return fail(c, "forward walking: estimator index out of range");
  const int kind = c->est[est_index].kind;
  if (kind == QMC_EST_STRUCTUREFACTOR)
    /* estimatorsFactory.cpp:66-69: only scalar and histogram observables have forward-walking estimators */
    return fail(c, "Forward walking estimator for vector not yet implemented");
  if (record_steps < 1 || n_fw < 1 || !fw_steps) return fail(c, "forward walking: recordSteps and forwardWalkingSteps are required");
  for (int k = 0; k < n_fw; ++k)
    if (fw_steps[k] < 0 || fw_steps[k] >= record_steps)
      return fail(c, "forward walking: forwardWalkingSteps must be in [0, recordSteps)");   /* assert, correlationEstimator.cpp:137 */
  if (first_index) *first_index = QMC_FW_INDEX_BASE + (int)c->fw_est.size();
  for (int k = 0; k < n_fw; ++k)
    {
      qmc_ctx::FwEst f;
      f.target = est_index;
      f.fw = fw_steps[k];
      f.record_steps = record_steps;
      c->fw_est.push_back(f);
    }
  return 0;
}

int qmc_enable_force_energy(qmc_ctx *c)
{
  if (!c) return 1;
  if (c->planned) return fail(c, "forceEnergy must be declared before qmc_set_dmc / qmc_set_vmc");
  c->force_energy = true;
  return 0;
}

int qmc_get_force_energy(qmc_ctx *c, double *sum, double *count)
{
  if (!c) return 1;
  if (!c->planned) return fail(c, "qmc_set_dmc must be called first");
  if (c->planning_only) return fail(c, "planning-only context (device -1): no CUDA device, no compute (there is no CPU fallback)");
  if (!c->force_energy) return fail(c, "forceEnergy was not enabled (qmc_enable_force_energy)");
  cudaSetDevice(c->device);
  double h[2] = {0, 0};
  if (c->nranks > 1)
    { /* estimator::accumulateMPI: reduced over the ranks */
      if (ensure_scratch(c, 64, 0)) return 1;
      CK(cudaMemcpyAsync(c->scr, c->fe_acc, 16, cudaMemcpyDeviceToDevice, c->stream));
      NK(g_nccl.AllReduce(c->scr, c->scr, 2, NCCL_FLOAT64, NCCL_SUM, c->comm, c->stream));
      CK(cudaMemcpyAsync(h, c->scr, 16, cudaMemcpyDeviceToHost, c->stream));
    }
  else CK(cudaMemcpyAsync(h, c->fe_acc, 16, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (sum) *sum = h[0];
  if (count) *count = h[1];
  return 0;
}

int qmc_add_superfluid_fraction(qmc_ctx *c, int setA, int setB, int bins, int *est_index)
{
  if (!c) return 1;
  if (c->planned) return fail(c, "superfluidFraction must be declared before qmc_set_dmc");
  if (setA < 0 || setB < 0 || setA >= (int)c->N.size() || setB >= (int)c->N.size()) return fail(c, "superfluidFraction: set out of range");
  if (bins < 1) return fail(c, "superfluidFraction: bins >= 1");
  if (c->sfl_est.empty()) c->sfl_record_steps = bins;       /* createSuperfluidFractionStorer: the first measurement's "bins" (:376-383) */
  else if (bins != c->sfl_record_steps)
    /* the reference derives the record size as data.size()/bins (:333): any other value reads the ring with the wrong stride */
    return fail(c, "superfluidFraction: every measurement must use the same \"bins\" (it is the centre-of-mass storer's recordSteps)");
  qmc_ctx::SflEst f;
  f.setA = setA;
  f.setB = setB;
  f.bins = bins;
  if (est_index) *est_index = QMC_SFL_INDEX_BASE + (int)c->sfl_est.size();
  c->sfl_est.push_back(f);
  return 0;
}

int qmc_set_dmc(qmc_ctx *c, double timeStep, int64_t targetWalkers, int64_t deltaWalkers, uint64_t seed,
                int branching_on, int64_t capacity)
{
  if (!c) return 1;
  if (c->planned) return fail(c, "qmc_set_dmc called twice");
  if (c->genealogy() || c->force_energy) c->pipeline = 1;      /* records / forceEnergy ride on the fused pipeline */
  if (c->force_energy && branching_on)
    /* main.cpp attaches the forceEnergy estimator to vmcDriver (:318-335) and to dmcDriver only for svmc (:395-408) */
    return fail(c, "forceEnergy is measured by the methods without branching (vmc, svmc), as in the reference");
  if (!c->planning_only) cudaSetDevice(c->device);
  if (build_plan(c)) return 1;
  DmcParams &D = c->dmc;
  D.tau = timeStep;
  D.sqrt_tau = std::sqrt(timeStep);
  D.mean_walkers = (double)targetWalkers;
  D.delta_walkers = (double)deltaWalkers;
  /* branching.cpp:84-96 (size_t arithmetic on meanWalkers/deltaWalkers) */
  D.log_up = 1. / timeStep * std::log(targetWalkers * 1. / (double)(targetWalkers + deltaWalkers));
  D.log_down = (targetWalkers > deltaWalkers) ? 1. / timeStep * std::log(targetWalkers * 1. / (double)(targetWalkers - deltaWalkers)) : 0.0;
  D.seed_lo = (uint32_t)seed;
  D.seed_hi = (uint32_t)(seed >> 32);
  D.branching_on = branching_on ? 1 : 0;
  D.rank = c->rank;
  D.nranks = c->nranks;
  if (capacity <= 0) capacity = 2 * (targetWalkers + deltaWalkers) / std::max(1, c->nranks) + 64;
  c->cap = capacity;
  if (c->genealogy())
    { /* ring depths (host logic): histogram rows for max fw + 1 events, links and centre-of-mass records for max(that, bins) */
      int maxfw = 0;
      for (auto &f : c->fw_est) maxfw = std::max(maxfw, f.fw);
      c->fw_depth = maxfw + 1;
      c->fw_link_depth = std::max(c->fw_depth, c->sfl_record_steps);
    }
  if (c->planning_only)
    {
      c->cfg.grid = c->sms * 2;
      c->planned = true;
      return 0;
    }
  const size_t dofs = c->plan.dofs;
  for (int b = 0; b < 2; ++b)
    {
      CK(cudaMalloc(&c->buf[b].R, capacity * dofs * 8));
      CK(cudaMalloc(&c->buf[b].F, capacity * dofs * 8));
      CK(cudaMalloc(&c->buf[b].logpsi, capacity * 8));
      CK(cudaMalloc(&c->buf[b].E, capacity * 8));
    }
  CK(cudaMalloc(&c->parent, capacity * 4));
  CK(cudaMalloc(&c->n_desc, capacity * 4));
  if (c->nranks > 1)
    { /* migration slabs sized once for a quarter of the capacity: no cudaMalloc/cudaFree (device-wide
         synchronisation) between the ranks' NCCL calls of qmc_rebalance in the common case */
      c->slab_walkers = (int)(capacity / 4 + 1);
      c->slab_bytes = (size_t)c->slab_walkers * (2 * dofs + 4) * 8;
      CK(cudaMalloc(&c->slab_send, c->slab_bytes));
      CK(cudaMalloc(&c->slab_recv, c->slab_bytes));
    }
  if (c->pipeline == 2)
    {
      CK(cudaMalloc(&c->e_old, capacity * 8));
      CK(cudaMalloc(&c->flags, capacity * 4));
      CK(cudaMemset(c->flags, 0, capacity * 4));
    }
  CK(cudaMalloc(&c->d_checksum, 8));
  /* staging for qmc_download_walkers sized once for the whole capacity: growing it with the population meant a cudaFree +
     cudaMalloc of hundreds of MB inside the end-to-end path (measured: up to 190 ms per call) */
  if (ensure_scratch(c, (size_t)capacity * (dofs + 1) * 8, 0)) return 1;
  {
    const int nb = scan_max_blocks();
    CK(cudaMalloc(&c->scan_ws_i, (size_t)(2 + 3 * nb) * sizeof(int)));
    CK(cudaMemset(c->scan_ws_i, 0, (size_t)(2 + 3 * nb) * sizeof(int)));
    CK(cudaMalloc(&c->scan_ws_d, (size_t)nb * sizeof(double)));
    c->scan_ws.ticket = c->scan_ws_i;
    c->scan_ws.done = c->scan_ws_i + 1;
    c->scan_ws.total = c->scan_ws_i + 2;
    c->scan_ws.nacc = c->scan_ws_i + 2 + nb;
    c->scan_ws.flag = reinterpret_cast<unsigned *>(c->scan_ws_i + 2 + 2 * nb);
    c->scan_ws.esum = c->scan_ws_d;
  }
  c->cfg.grid = grid_for(c, capacity);
  {
    const Plan &P = c->plan;
    const char *le = getenv("QMC4_LEAN");
    c->lean = c->pipeline == 2 && !(le && atoi(le) == 0) && P.nSpecies == 1 && P.n1b == 0 && P.n2b == 1 && P.n3b == 0 &&
              P.p2[0].setA == 0 && P.p2[0].setB == 0 && P.p2[0].J.kind == J_GAUSSGAUSS && (P.p2[0].vkind == V_STEP || P.p2[0].vkind == V_NONE);
    if (c->lean)
      {
        c->lean_vk = P.p2[0].vkind;
        if (const char *o = getenv("QMC4_LEAN_VARIANT")) c->lean_occ = std::max(0, std::min(3, atoi(o)));
        LaunchCfg &L = c->lean_cfg;
        for (;;)
          {
            L.block = pair_variant_threads(c->lean_occ);
            L.nwarps_per_group = std::min(pair_warps_per_walker(P.N[0]), L.block / 32);
            L.groups_per_cta = (L.block / 32) / L.nwarps_per_group;
            L.smem = pair_group_bytes(P, L.nwarps_per_group) * L.groups_per_cta + pair_cta_extra_bytes(c->lean_occ);
            if ((int)L.smem <= c->max_smem || c->lean_occ != 3) break;
            c->lean_occ = 0;      /* 16 walker groups + the replicated table do not fit: two 256-thread CTAs with one table each */
          }
        if ((int)L.smem > c->max_smem) c->lean = false;
        else
          {
            int per_sm = std::max(1, std::min(pair_move_occupancy(c->lean_vk, c->lean_occ, L.smem), 2048 / L.block));
            per_sm = std::min(per_sm, c->lean_occ == 3 ? 1 : c->lean_occ == 2 ? 5 : c->lean_occ == 1 ? 4 : 2);
            if (const char *e2 = getenv("QMC4_CTAS_PER_SM"))
              {
                const int v = atoi(e2);
                if (v >= 1 && v <= 8) per_sm = v;
              }
            const int64_t want = (capacity + L.groups_per_cta - 1) / L.groups_per_cta;
            L.grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)compute_sms(c) * per_sm));
            double tab[256];
            pair_exp_table(tab);
            CK(cudaMalloc(&c->lean_tab, sizeof tab));
            CK(cudaMemcpy(c->lean_tab, tab, sizeof tab, cudaMemcpyHostToDevice));
          }
      }
  }
  {
    const Plan &P = c->plan;
    const char *se = getenv("QMC4_SMALL");
    bool ok = c->pipeline == 2 && !c->lean && !(se && atoi(se) == 0) && P.nSpecies == 1 && P.N[0] <= 16 && P.n3b == 0;
    for (int t = 0; t < P.n2b; ++t) ok = ok && P.p2[t].setA == 0 && P.p2[t].setB == 0;
    c->small = ok;
    if (ok)
      {
        const int wpc = small_walkers_per_cta(P.N[0]);
        const int64_t want = (capacity + wpc - 1) / wpc;
        c->small_grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)compute_sms(c) * small_move_occupancy(P.N[0])));
      }
  }
  c->n_partials = c->cfg.grid * c->cfg.groups_per_cta;
  CK(cudaMalloc(&c->partials, (size_t)c->n_partials * 2 * 8));
  CK(cudaMemset(c->partials, 0, (size_t)c->n_partials * 2 * 8));
  if (!c->coefpool.empty())
    {
      CK(cudaMalloc(&c->coef, c->coefpool.size() * 8));
      CK(cudaMemcpy(c->coef, c->coefpool.data(), c->coefpool.size() * 8, cudaMemcpyHostToDevice));
    }
  if (c->trip_cyclic)
    { /* canonical offset pairs (p, q) modulo M: gaps (p, q-p, M-q) equal to their lexicographically smallest rotation */
      const int M = c->trip_mod;
      std::vector<int> tabv;
      for (int p = 1; p < M; ++p)
        for (int q = p + 1; q < M; ++q)
          {
            const int g[3] = {p, q - p, M - q};
            bool canon = true, sym = true;
            for (int r = 1; r < 3; ++r)
              {
                const int h[3] = {g[r], g[(r + 1) % 3], g[(r + 2) % 3]};
                if (std::lexicographical_compare(h, h + 3, g, g + 3)) canon = false;
                if (h[0] != g[0] || h[1] != g[1] || h[2] != g[2]) sym = false;
              }
            if (canon) tabv.push_back(p | (q << 8) | ((sym ? 1 : 0) << 16));
          }
      CK(cudaMalloc(&c->trip_tab, tabv.size() * sizeof(int)));
      CK(cudaMemcpy(c->trip_tab, tabv.data(), tabv.size() * sizeof(int), cudaMemcpyHostToDevice));
      c->plan.trip_tab = c->trip_tab;
      c->plan.trip_count = (int)tabv.size();
      c->plan.trip_mod = M;
    }
  if (c->plan.hist_stride > 0)
    {
      c->plan.hist_replicas = c->cfg.grid;
      c->est_grid = c->cfg.grid;
      if (c->pipeline == 2)
        { /* the estimator kernel of pipeline 2 has its own occupancy target (QMC4_EST_OCC = 2|3|4 resident CTAs per SM) */
          if (c->max_sf_bins > 0) c->est_occ = 2;      /* the structure factor carries 8 phasors per thread in registers */
          if (const char *e = getenv("QMC4_EST_OCC"))
            {
              const int v = atoi(e);
              if (v >= 2 && v <= 4) c->est_occ = v;
            }
          const int per_sm = std::max(1, std::min(estimator_occupancy(c->est_occ, c->cfg.smem), c->est_occ));
          const int64_t want = (capacity + c->cfg.groups_per_cta - 1) / c->cfg.groups_per_cta;
          c->est_grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)compute_sms(c) * per_sm));
          c->plan.hist_replicas = std::max(c->cfg.grid, c->est_grid);
        }
      {
        const Plan &P = c->plan;
        const char *he = getenv("QMC4_LEAN_HIST");
        c->lean_hist = c->pipeline == 2 && c->lean && !(he && atoi(he) == 0) && P.nest == 1 && P.est[0].kind == QMC_EST_PAIRCORRELATION &&
                       P.est[0].setA == 0 && P.est[0].setB == 0 && (int)pair_hist_smem(P) <= c->max_smem;
        if (c->lean_hist)
          {
            const int64_t want = (capacity + 7) / 8;
            c->lean_hist_grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)compute_sms(c) * std::min(4, pair_hist_occupancy(P))));
            c->plan.hist_replicas = std::max(c->plan.hist_replicas, c->lean_hist_grid);
          }
      }
      CK(cudaMalloc(&c->hist, (size_t)c->plan.hist_replicas * c->plan.hist_stride * 8));
      CK(cudaMemset(c->hist, 0, (size_t)c->plan.hist_replicas * c->plan.hist_stride * 8));
      CK(cudaMalloc(&c->hist_out, (size_t)c->plan.hist_stride * 8));
    }
  if (c->force_energy)
    {
      CK(cudaMalloc(&c->fe_R, capacity * c->plan.dofs * 8));
      CK(cudaMalloc(&c->fe_F, capacity * c->plan.dofs * 8));
      CK(cudaMalloc(&c->fe_V, capacity * 8));
      CK(cudaMalloc(&c->fe_partial, FE_BLOCKS * 8));
      CK(cudaMalloc(&c->fe_acc, 16));
      CK(cudaMemset(c->fe_acc, 0, 16));
    }
  if (c->genealogy())
    { /* the event clock both record rings share: ancestor links + the composed parent maps between sampling events */
      CK(cudaMalloc(&c->fw_links, (size_t)c->fw_link_depth * capacity * 4));
      CK(cudaMalloc(&c->fw_comp[0], capacity * 4));
      CK(cudaMalloc(&c->fw_comp[1], capacity * 4));
      CK(cudaMalloc(&c->fw_level_slots, (size_t)c->fw_link_depth * 4));
      CK(cudaMemset(c->fw_level_slots, 0, (size_t)c->fw_link_depth * 4));
    }
  if (!c->sfl_est.empty())
    {
      const size_t width = 3 * (size_t)c->plan.nSpecies, R = c->sfl_record_steps;
      const size_t ring = (size_t)c->fw_link_depth * capacity * width * 8;
      if (cudaMalloc(&c->sfl_rec, ring) != cudaSuccess)
        return fail(c, "superfluidFraction: cannot allocate the centre-of-mass ring (" + std::to_string(ring >> 20) +
                           " MiB = bins x capacity x 3 nSpecies x 8 B): fewer walkers or bins");
      const size_t chunks = (capacity + sfl_chunk_slots() - 1) / sfl_chunk_slots();
      CK(cudaMalloc(&c->sfl_partial, chunks * R * 8));
      for (auto &f : c->sfl_est)
        {
          CK(cudaMalloc(&f.sums, R * 8));
          CK(cudaMalloc(&f.weights, R * 8));
          CK(cudaMemset(f.sums, 0, R * 8));
          CK(cudaMemset(f.weights, 0, R * 8));
        }
    }
  if (!c->fw_est.empty())
    {
      int maxbins = 1;
      for (auto &f : c->fw_est) maxbins = std::max(maxbins, c->plan.est[f.target].bins);
      CK(configure_kernels_fw(c->max_smem));
      c->fw_group_bytes = fw_record_group_bytes(c->plan, c->max_sf_bins);
      if (c->fw_group_bytes + (size_t)c->plan.hist_stride * 8 > (size_t)c->max_smem)
        return fail(c, "forward walking: the record kernel needs more shared memory than one CTA has for this system size");
      const size_t ring = (size_t)c->fw_depth * capacity * c->plan.hist_stride * 8;
      if (cudaMalloc(&c->fw_rows, ring) != cudaSuccess)
        return fail(c, "forward walking: cannot allocate the record ring (" + std::to_string(ring >> 20) +
                           " MiB = (max forwardWalkingSteps+1) x capacity x bins x 8 B): fewer walkers, bins or forward-walking steps");
      CK(cudaMalloc(&c->fw_cnt, capacity * 4));
      CK(cudaMalloc(&c->fw_nslots, (size_t)c->fw_depth * 4));
      CK(cudaMemset(c->fw_nslots, 0, (size_t)c->fw_depth * 4));
      CK(cudaMalloc(&c->fw_partial, (size_t)fw_chunks() * maxbins * 8));
      for (auto &f : c->fw_est)
        {
          const int bins = c->plan.est[f.target].bins;
          CK(cudaMalloc(&f.sums, (size_t)bins * 8));
          CK(cudaMalloc(&f.weight, 8));
          CK(cudaMemset(f.sums, 0, (size_t)bins * 8));
          CK(cudaMemset(f.weight, 0, 8));
        }
    }
  c->planned = true;
  return 0;
}

int qmc_tabulate_wavefunction(qmc_ctx *c, int wf_index, double minx, double maxx, int64_t n, double *out)
{
  if (!c) return 1;
  if (!c->planned) return fail(c, "qmc_set_dmc must be called first");
  if (c->planning_only) return fail(c, "planning-only context (device -1): no CUDA device, no compute (there is no CPU fallback)");
  if (wf_index < 0 || wf_index >= (int)c->wf.size()) return fail(c, "qmc_tabulate_wavefunction: term index out of range");
  if (n < 1 || !out) return fail(c, "qmc_tabulate_wavefunction: n >= 1 and an output buffer are required");   /* assert(n>=1), jastrow.h:45 */
  cudaSetDevice(c->device);
  if (ensure_scratch(c, (size_t)n * 4 * 8, 0)) return 1;
  CK(launch_tabulate(c->stream, c->wf[wf_index].J, c->coef, minx, (maxx - minx) / (double)n, (long)n, c->scr));
  c->launches += 1;
  CK(cudaMemcpyAsync(out, c->scr, (size_t)n * 4 * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

int qmc_set_vmc(qmc_ctx *c, double sigma, int64_t nChains, uint64_t seed, int64_t capacity)
{
  if (!c) return 1;
  if (!(sigma > 0) || nChains < 1) return fail(c, "qmc_set_vmc: sigma > 0 and at least one chain");
  if (!c->fw_est.empty()) return fail(c, "forward walking needs branching genealogies: use method dmc");
  if (!c->sfl_est.empty()) return fail(c, "Cannot have vmc and superfluid fraction measurements");        /* main.cpp:279-282 */
  c->pipeline = 1;                      /* the Green-ratio switch lives in propagate_kernel */
  if (capacity <= 0) capacity = nChains / std::max(1, c->nranks) + 64;
  if (qmc_set_dmc(c, sigma, nChains, 0, seed, 0, capacity)) return 1;
  DmcParams &D = c->dmc;
  D.tau = 0.0;                          /* no drift:  x' = x + xi*sqrt(sigma)  (vmcMoves.cpp:31) */
  D.sqrt_tau = std::sqrt(sigma);
  D.log_up = D.log_down = 0.0;
  D.vmc = 1;
  return 0;
}

static int eval_common(qmc_ctx *c, const double *dR, int64_t nW, double *dlogpsi, double *dgrad, double *dlap,
                       double *dV, double *dE, int *dvalid, bool hist, int flag_invalid)
{
  EvalArgs a;
  memset(&a, 0, sizeof a);
  a.R = dR;
  a.nW = (int)nW;
  a.logpsi = dlogpsi;
  a.F = dgrad;
  a.lap = dlap;
  a.V = dV;
  a.E = dE;
  a.valid = dvalid;
  a.coef = c->coef;
  a.hist = hist ? c->hist : nullptr;
  a.ctl = c->ctl;
  a.flag_invalid = flag_invalid;
  LaunchCfg cfg = c->cfg;
  cfg.grid = std::min(c->cfg.grid, grid_for(c, nW));
  CK(launch_eval(c->ksel, hist, cfg, c->stream, c->plan, a));
  c->launches += 1;
  return 0;
}

int qmc_eval_device(qmc_ctx *c, const double *dR, int64_t nW, double *dlogpsi, double *dgrad, double *dlap, double *dV,
                    double *dE, int *dvalid)
{
  if (!c) return 1;
  if (!c->planned) return fail(c, "qmc_set_dmc must be called first");
  if (c->planning_only) return fail(c, "planning-only context (device -1): no CUDA device, no compute (there is no CPU fallback)");
  cudaSetDevice(c->device);
  if (nW <= 0) return 0;
  return eval_common(c, dR, nW, dlogpsi, dgrad, dlap, dV, dE, dvalid, false, 0);
}

int qmc_eval(qmc_ctx *c, const double *R, int64_t nW, double *logpsi, double *grad, double *lap, double *V, double *E, int *valid)
{
  if (!c) return 1;
  if (!c->planned) return fail(c, "qmc_set_dmc must be called first");
  if (c->planning_only) return fail(c, "planning-only context (device -1): no CUDA device, no compute (there is no CPU fallback)");
  cudaSetDevice(c->device);
  if (nW <= 0) return 0;
  const size_t dofs = c->plan.dofs;
  const size_t need = (size_t)nW * (2 * dofs + 4) * 8;
  if (ensure_scratch(c, need, (size_t)nW * 4)) return 1;
  double *dR = c->scr, *dF = dR + nW * dofs, *dlp = dF + nW * dofs, *dlap = dlp + nW, *dV = dlap + nW, *dE = dV + nW;
  CK(cudaMemcpyAsync(dR, R, nW * dofs * 8, cudaMemcpyHostToDevice, c->stream));
  if (eval_common(c, dR, nW, dlp, dF, dlap, dV, dE, c->scr_i, false, 0)) return 1;
  if (logpsi) CK(cudaMemcpyAsync(logpsi, dlp, nW * 8, cudaMemcpyDeviceToHost, c->stream));
  if (grad) CK(cudaMemcpyAsync(grad, dF, nW * dofs * 8, cudaMemcpyDeviceToHost, c->stream));
  if (lap) CK(cudaMemcpyAsync(lap, dlap, nW * 8, cudaMemcpyDeviceToHost, c->stream));
  if (V) CK(cudaMemcpyAsync(V, dV, nW * 8, cudaMemcpyDeviceToHost, c->stream));
  if (E) CK(cudaMemcpyAsync(E, dE, nW * 8, cudaMemcpyDeviceToHost, c->stream));
  if (valid) CK(cudaMemcpyAsync(valid, c->scr_i, nW * 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

int qmc_accumulate_configs(qmc_ctx *c, const double *R, int64_t nW)
{
  if (!c) return 1;
  if (!c->planned) return fail(c, "qmc_set_dmc must be called first");
  if (c->planning_only) return fail(c, "planning-only context (device -1): no CUDA device, no compute (there is no CPU fallback)");
  if (c->pipeline == 2 && sync_all(c)) return 1;
  if (!c->has_hist) return fail(c, "no estimators registered");
  cudaSetDevice(c->device);
  if (nW <= 0) return 0;
  const size_t dofs = c->plan.dofs;
  if (ensure_scratch(c, (size_t)nW * (dofs + 1) * 8, (size_t)nW * 4)) return 1;
  CK(cudaMemcpyAsync(c->scr, R, nW * dofs * 8, cudaMemcpyHostToDevice, c->stream));
  if (eval_common(c, c->scr, nW, nullptr, nullptr, nullptr, nullptr, nullptr, c->scr_i, true, 0)) return 1;
  std::vector<int> valid(nW);
  CK(cudaMemcpyAsync(valid.data(), c->scr_i, nW * 4, cudaMemcpyDeviceToHost, c->stream));
  Ctl h;
  CK(cudaMemcpyAsync(&h, c->ctl, sizeof h, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  double w = 0;
  for (int v : valid) w += v ? 1 : 0;
  h.est_weight += w;
  CK(cudaMemcpy(&c->ctl->est_weight, &h.est_weight, 8, cudaMemcpyHostToDevice));
  return 0;
}

int qmc_upload_walkers(qmc_ctx *c, const double *R, int64_t nW)
{
  if (!c) return 1;
  if (!c->planned) return fail(c, "qmc_set_dmc must be called first");
  if (c->planning_only) return fail(c, "planning-only context (device -1): no CUDA device, no compute (there is no CPU fallback)");
  if (c->pipeline == 2 && sync_all(c)) return 1;
  cudaSetDevice(c->device);
  if (nW < 1) return fail(c, "need at least one walker");
  if (nW > c->cap) return fail(c, "more walkers than capacity");
  c->fw_event = 0;                     /* a new population starts with empty records (storer::reserve) */
  c->n_host = nW;
  const size_t dofs = c->plan.dofs;
  WalkerBuf &B = c->buf[c->cur];
  CK(cudaMemcpyAsync(B.R, R, nW * dofs * 8, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemsetAsync(&c->ctl->invalid_init, 0, 4, c->stream));
  if (eval_common(c, B.R, nW, B.logpsi, B.F, nullptr, nullptr, B.E, nullptr, false, 1)) return 1;
  CK(launch_init_identity(c->stream, c->parent, c->n_desc, (int)nW, c->ctl));
  // first E_T: brancher->setEnergyShift(current_walkers) (dmcDriver.cpp:191)
  StepArgs a = make_step_args(c, 0);
  a.out = B;
  CK(launch_scan(c->stream, c->dmc, a, 0, c->nranks > 1, 1));
  c->launches += 2;
  if (allreduce_sums(c)) return 1;
  if (c->nranks > 1 && !c->p2p_connected)
    { /* NCCL sets up point-to-point connections lazily, one directed pair at a time (~140 ms each on NVLink, measured): a
         migration inside a step must never pay that, so every pair exchanges one double here, once, in a single group */
      NK(g_nccl.GroupStart());
      for (int r = 0; r < c->nranks; ++r)
        if (r != c->rank)
          {
            NK(g_nccl.Send(c->slab_send + r, 1, NCCL_FLOAT64, r, c->comm, c->stream));
            NK(g_nccl.Recv(c->slab_recv + r, 1, NCCL_FLOAT64, r, c->comm, c->stream));
          }
      NK(g_nccl.GroupEnd());
      c->p2p_connected = true;
    }
  Ctl h;
  CK(cudaMemcpyAsync(&h, c->ctl, sizeof h, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (c->nranks > 1)
    { /* the migration planner of the first step reads these populations; every rank must own the same number of slots
         (the plan is clamped to the receivers' free slots and to the slab size, identically on every rank) */
      memcpy(c->h_sums, h.gsum, sizeof(double) * NSUM);
      c->pops_pending = true;
      c->ar_pending = false;
      for (int r = 1; r < c->nranks; ++r)      /* free slots + population = capacity of rank r */
        if (h.gsum[2 + MAX_RANKS + r] + h.gsum[2 + r] != h.gsum[2 + MAX_RANKS] + h.gsum[2])
          return fail(c, "every rank must be given the same capacity (qmc_set_dmc)");
    }
  if (h.invalid_init) return fail(c, "Inalid initial configurations.");   /* dmcDriver.cpp:166-170 (sic) */
  return 0;
}

int qmc_download_walkers(qmc_ctx *c, double *R, int64_t capW, int64_t *nW, double *E)
{
  if (!c) return 1;
  if (!c->planned) return fail(c, "qmc_set_dmc must be called first");
  if (c->planning_only) return fail(c, "planning-only context (device -1): no CUDA device, no compute (there is no CPU fallback)");
  if (c->pipeline == 2 && sync_all(c)) return 1;
  cudaSetDevice(c->device);
  Ctl h;
  CK(cudaMemcpyAsync(&h, c->ctl, sizeof h, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  const int64_t n = h.n_cur;
  if (nW) *nW = n;
  if (!R || n == 0) return 0;
  if (n > capW) return fail(c, "download buffer too small");
  const size_t dofs = c->plan.dofs;
  if (ensure_scratch(c, (size_t)n * (dofs + 1) * 8, 0)) return 1;
  QMC_LAUNCH(gather_kernel, (int)n, 128, 0, c->stream, c->buf[c->cur], c->parent, (int)dofs, (int)n, c->scr, c->scr + n * dofs);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(R, c->scr, n * dofs * 8, cudaMemcpyDeviceToHost, c->stream));
  if (E) CK(cudaMemcpyAsync(E, c->scr + n * dofs, n * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

int qmc_step(qmc_ctx *c, int nSteps, int sampleEvery)
{
  if (!c) return 1;
  if (!c->planned) return fail(c, "qmc_set_dmc must be called first");
  if (c->planning_only) return fail(c, "planning-only context (device -1): no CUDA device, no compute (there is no CPU fallback)");
  cudaSetDevice(c->device);
  if (sampleEvery < 1) sampleEvery = 1;
  const double h0 = c->trace_n ? host_now_ms() : 0.0;
  c->trace_used = 0;
  CK(cudaEventRecord(c->ev0, c->stream));
  int k0 = 0;
#ifndef QMC_EMULATE
  if (c->use_graphs && c->pipeline == 2 && c->nranks == 1 && sampleEvery == 1 && !c->time_kernels && !c->trace_n && !c->genealogy() &&
      !c->force_energy && nSteps >= 4)
    { /* launch-bound regime (config 1: two 10 us kernels per step): replay a captured pair of steps */
      if (c->cur == 1)
        { /* the graph was captured from buffer 0 */
          c->sample_phase += 1;
          if (one_step(c, 1, nullptr)) return 1;
          k0 = 1;
        }
      if (!c->step_graph)
        {
          cudaGraph_t g = nullptr;
          const long l0 = c->launches;
          CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
          int rc = one_step(c, 1, nullptr);
          if (!rc) rc = one_step(c, 1, nullptr);
          cudaError_t e = cudaStreamEndCapture(c->stream, &g);
          if (rc || e != cudaSuccess) return fail(c, std::string("graph capture of the step pair failed: ") + (rc ? c->err : cudaGetErrorString(e)));
          c->graph_launches_per_replay = c->launches - l0;
          c->launches = l0;                      /* nothing ran during the capture */
          CK(cudaGraphInstantiate(&c->step_graph, g, 0));
          cudaGraphDestroy(g);
        }
      for (; k0 + 2 <= nSteps; k0 += 2)
        {
          CK(cudaGraphLaunch(c->step_graph, c->stream));
          c->sample_phase += 2;
          c->launches += c->graph_launches_per_replay;
          c->graph_replays += 1;
        }
    }
#endif
  for (int k = k0; k < nSteps; ++k)
    {
      c->sample_phase += 1;
      const int accumulate = (c->sample_phase % sampleEvery) == 0;
      if (one_step(c, accumulate, nullptr)) return 1;
    }
  CK(cudaEventRecord(c->ev1, c->stream));
  CK(cudaEventSynchronize(c->ev1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  c->device_ms += ms;
  if (c->trace_n) dump_trace(c, h0);
  return resolve_kernel_timers(c);
}

int qmc_step_traced(qmc_ctx *c, int accumulate, const qmc_trace *tr, int64_t traceCap, int64_t *nW_before, int64_t *nW_after)
{
  if (!c) return 1;
  if (!c->planned) return fail(c, "qmc_set_dmc must be called first");
  if (c->planning_only) return fail(c, "planning-only context (device -1): no CUDA device, no compute (there is no CPU fallback)");
  if (c->pipeline == 2 && sync_all(c)) return 1;
  cudaSetDevice(c->device);
  Ctl h;
  CK(cudaMemcpyAsync(&h, c->ctl, sizeof h, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  const int64_t n = h.n_cur;
  if (nW_before) *nW_before = n;
  if (tr && n > traceCap) return fail(c, "trace buffers too small");
  const size_t dofs = c->plan.dofs;
  TraceDev T;
  memset(&T, 0, sizeof T);
  double *dbuf = nullptr;
  int *ibuf = nullptr;
  if (tr)
    {
      CK(cudaMalloc(&dbuf, (size_t)n * (3 * dofs + 7) * 8));
      CK(cudaMalloc(&ibuf, (size_t)n * 4 * 4));
      T.xi = dbuf;
      T.Rprop = T.xi + n * dofs;
      T.Fprop = T.Rprop + n * dofs;
      T.logpsi = T.Fprop + n * dofs;
      T.lap = T.logpsi + n;
      T.V = T.lap + n;
      T.Eprop = T.V + n;
      T.q = T.Eprop + n;
      T.u_m = T.q + n;
      T.u_b = T.u_m + n;
      T.accepted = ibuf;
      T.valid = ibuf + n;
      T.n_desc = ibuf + 2 * n;
      T.parent = ibuf + 3 * n;
    }
  int rc = one_step(c, accumulate, tr ? &T : nullptr);
  if (!rc && tr)
    {
      auto D = [&](void *dst, const void *src, size_t bytes) {
        if (dst) cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream);
      };
      D(tr->xi, T.xi, n * dofs * 8);
      D(tr->Rprop, T.Rprop, n * dofs * 8);
      D(tr->Fprop, T.Fprop, n * dofs * 8);
      D(tr->logpsi, T.logpsi, n * 8);
      D(tr->lap, T.lap, n * 8);
      D(tr->V, T.V, n * 8);
      D(tr->Eprop, T.Eprop, n * 8);
      D(tr->q, T.q, n * 8);
      D(tr->u_metropolis, T.u_m, n * 8);
      D(tr->u_branch, T.u_b, n * 8);
      D(tr->accepted, T.accepted, n * 4);
      D(tr->valid, T.valid, n * 4);
      D(tr->n_desc, T.n_desc, n * 4);
      D(tr->parent, T.parent, n * 4);
    }
  cudaStreamSynchronize(c->stream);
  if (dbuf) cudaFree(dbuf);
  if (ibuf) cudaFree(ibuf);
  if (rc) return rc;
  CK(cudaMemcpy(&h, c->ctl, sizeof h, cudaMemcpyDeviceToHost));
  if (nW_after) *nW_after = h.n_cur;
  return 0;
}

int qmc_block_results(qmc_ctx *c, qmc_block_result *out)
{
  if (!c || !out) return 1;
  if (!c->planned) return fail(c, "qmc_set_dmc must be called first");
  if (c->planning_only) return fail(c, "planning-only context (device -1): no CUDA device, no compute (there is no CPU fallback)");
  if (c->pipeline == 2 && sync_all(c)) return 1;
  cudaSetDevice(c->device);
  Ctl h;
  CK(cudaMemcpyAsync(&h, c->ctl, sizeof h, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  /* the overflow flag travels with the sums: every rank sees it, so the host driver's error is collective too */
  double v[5] = {h.blk_esum, h.blk_ecount, h.blk_moves, h.blk_acc, (double)(h.overflow != 0)};
  if (c->nranks > 1)
    {
      if (ensure_scratch(c, 64, 0)) return 1;
      CK(cudaMemcpyAsync(c->scr, v, 40, cudaMemcpyHostToDevice, c->stream));
      NK(g_nccl.AllReduce(c->scr, c->scr, 5, NCCL_FLOAT64, NCCL_SUM, c->comm, c->stream));
      CK(cudaMemcpyAsync(v, c->scr, 40, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
    }
  out->energy_sum = v[0];
  out->energy_count = v[1];
  out->n_moves = v[2];
  out->n_accepted = v[3];
  out->population = h.n_cur;
  {
    const DmcParams &D = c->dmc;
    double e = h.gsum[1] > 0 ? h.gsum[0] / h.gsum[1] : 0.0;
    if (h.gsum[1] > D.mean_walkers + D.delta_walkers) e += D.log_up;
    else if (h.gsum[1] < D.mean_walkers - D.delta_walkers) e += D.log_down;
    out->energy_shift = D.branching_on ? e : 0.0;
  }
  out->steps_done = h.blk_steps;
  out->overflow = v[4] != 0.0;
  out->device_ms = c->device_ms;
  c->device_ms = 0;
  CK(cudaMemsetAsync(&c->ctl->blk_esum, 0, 5 * 8, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

int qmc_get_estimator(qmc_ctx *c, int est_index, double *sums, int bins, double *weight)
{
  if (!c) return 1;
  if (!c->planned) return fail(c, "qmc_set_dmc must be called first");
  if (c->planning_only) return fail(c, "planning-only context (device -1): no CUDA device, no compute (there is no CPU fallback)");
  if (est_index >= QMC_FW_INDEX_BASE)
    { /* forward-walking estimator: sums and weight of realHistogramForwardWalkingEstimator */
      const int k = est_index - QMC_FW_INDEX_BASE;
      if (k >= (int)c->fw_est.size()) return fail(c, "estimator index out of range");
      const auto &f = c->fw_est[k];
      if (bins != c->plan.est[f.target].bins) return fail(c, "bins mismatch");
      cudaSetDevice(c->device);
      double w = 0;
      if (c->nranks > 1)
        { /* estimator::accumulateMPI: sums and weight reduced over the ranks */
          if (ensure_scratch(c, (size_t)(bins + 1) * 8, 0)) return 1;
          CK(cudaMemcpyAsync(c->scr, f.sums, (size_t)bins * 8, cudaMemcpyDeviceToDevice, c->stream));
          CK(cudaMemcpyAsync(c->scr + bins, f.weight, 8, cudaMemcpyDeviceToDevice, c->stream));
          NK(g_nccl.AllReduce(c->scr, c->scr, (size_t)bins + 1, NCCL_FLOAT64, NCCL_SUM, c->comm, c->stream));
          CK(cudaMemcpyAsync(sums, c->scr, (size_t)bins * 8, cudaMemcpyDeviceToHost, c->stream));
          CK(cudaMemcpyAsync(&w, c->scr + bins, 8, cudaMemcpyDeviceToHost, c->stream));
        }
      else
        {
          CK(cudaMemcpyAsync(sums, f.sums, (size_t)bins * 8, cudaMemcpyDeviceToHost, c->stream));
          CK(cudaMemcpyAsync(&w, f.weight, 8, cudaMemcpyDeviceToHost, c->stream));
        }
      CK(cudaStreamSynchronize(c->stream));
      if (weight) *weight = w;
      return 0;
    }
  if (est_index < 0 || est_index >= c->plan.nest) return fail(c, "estimator index out of range");
  if (c->pipeline == 2 && sync_all(c)) return 1;
  const Estimator &E = c->plan.est[est_index];
  if (bins != E.bins) return fail(c, "bins mismatch");
  cudaSetDevice(c->device);
  CK(launch_reduce_hist(c->stream, c->hist, c->plan.hist_replicas, c->plan.hist_stride, c->hist_out));
  c->launches += 1;
  Ctl h;
  CK(cudaMemcpyAsync(&h, c->ctl, sizeof h, cudaMemcpyDeviceToHost, c->stream));
  double w = 0;
  if (c->nranks > 1)
    {
      NK(g_nccl.AllReduce(c->hist_out, c->hist_out, c->plan.hist_stride, NCCL_FLOAT64, NCCL_SUM, c->comm, c->stream));
      if (ensure_scratch(c, 64, 0)) return 1;
      CK(cudaStreamSynchronize(c->stream));
      w = h.est_weight;
      CK(cudaMemcpyAsync(c->scr, &w, 8, cudaMemcpyHostToDevice, c->stream));
      NK(g_nccl.AllReduce(c->scr, c->scr, 1, NCCL_FLOAT64, NCCL_SUM, c->comm, c->stream));
      CK(cudaMemcpyAsync(&w, c->scr, 8, cudaMemcpyDeviceToHost, c->stream));
    }
  CK(cudaMemcpyAsync(sums, c->hist_out + E.offset, (size_t)bins * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (c->nranks == 1) w = h.est_weight;
  if (weight) *weight = w;
  return 0;
}

int qmc_clear_estimators(qmc_ctx *c)
{
  if (!c) return 1;
  if (!c->planned) return fail(c, "qmc_set_dmc must be called first");
  if (c->planning_only) return fail(c, "planning-only context (device -1): no CUDA device, no compute (there is no CPU fallback)");
  if (c->pipeline == 2 && sync_all(c)) return 1;
  cudaSetDevice(c->device);
  if (c->hist) CK(cudaMemsetAsync(c->hist, 0, (size_t)c->plan.hist_replicas * c->plan.hist_stride * 8, c->stream));
  CK(cudaMemsetAsync(&c->ctl->est_weight, 0, 8, c->stream));
  for (auto &f : c->fw_est)
    { /* accumulators are cleared every block (estimatorCollection::clear); the walkers' rings are not */
      CK(cudaMemsetAsync(f.sums, 0, (size_t)c->plan.est[f.target].bins * 8, c->stream));
      CK(cudaMemsetAsync(f.weight, 0, 8, c->stream));
    }
  if (c->fe_acc) CK(cudaMemsetAsync(c->fe_acc, 0, 16, c->stream));
  for (auto &f : c->sfl_est)
    {
      CK(cudaMemsetAsync(f.sums, 0, (size_t)f.bins * 8, c->stream));
      CK(cudaMemsetAsync(f.weights, 0, (size_t)f.bins * 8, c->stream));
    }
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

/* superfluidFractionEstimator's realVectorAccumulator: sums[bins] and weights[bins] (accumulators.h, correlationEstimator.cpp:305-320) */
int qmc_get_vector_estimator(qmc_ctx *c, int est_index, double *sums, double *weights, int bins)
{
  if (!c) return 1;
  if (!c->planned) return fail(c, "qmc_set_dmc must be called first");
  if (c->planning_only) return fail(c, "planning-only context (device -1): no CUDA device, no compute (there is no CPU fallback)");
  const int k = est_index - QMC_SFL_INDEX_BASE;
  if (k < 0 || k >= (int)c->sfl_est.size()) return fail(c, "estimator index out of range");
  const auto &f = c->sfl_est[k];
  if (bins != f.bins || !sums || !weights) return fail(c, "bins mismatch");
  cudaSetDevice(c->device);
  if (c->nranks > 1)
    {
      if (ensure_scratch(c, (size_t)bins * 16, 0)) return 1;
      CK(cudaMemcpyAsync(c->scr, f.sums, (size_t)bins * 8, cudaMemcpyDeviceToDevice, c->stream));
      CK(cudaMemcpyAsync(c->scr + bins, f.weights, (size_t)bins * 8, cudaMemcpyDeviceToDevice, c->stream));
      NK(g_nccl.AllReduce(c->scr, c->scr, (size_t)bins * 2, NCCL_FLOAT64, NCCL_SUM, c->comm, c->stream));
      CK(cudaMemcpyAsync(sums, c->scr, (size_t)bins * 8, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaMemcpyAsync(weights, c->scr + bins, (size_t)bins * 8, cudaMemcpyDeviceToHost, c->stream));
    }
  else
    {
      CK(cudaMemcpyAsync(sums, f.sums, (size_t)bins * 8, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaMemcpyAsync(weights, f.weights, (size_t)bins * 8, cudaMemcpyDeviceToHost, c->stream));
    }
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

int qmc_synchronize(qmc_ctx *c)
{
  if (!c) return 1;
  if (c->planning_only) return 0;
  cudaSetDevice(c->device);
  if (c->pipeline == 2 && sync_all(c)) return 1;
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

// ---- multi-GPU -----------------------------------------------------------------------------------
int qmc_comm_unique_id(void *uniqueId128)
{
  std::string err;
  if (!g_nccl.load(err))
    {
      g_create_error = err;
      return 1;
    }
  ncclUniqueId_t id;
  int e = g_nccl.GetUniqueId(&id);
  if (e)
    {
      g_create_error = g_nccl.GetErrorString(e);
      return 1;
    }
  memcpy(uniqueId128, &id, 128);
  return 0;
}

int qmc_comm_init(qmc_ctx *c, const void *uniqueId128, int rank, int nranks)
{
  if (!c) return 1;
  if (c->planning_only) return fail(c, "planning-only context (device -1): no communicator");
  if (c->planned) return fail(c, "qmc_comm_init must precede qmc_set_dmc");
  if (nranks < 1 || nranks > MAX_RANKS || rank < 0 || rank >= nranks) return fail(c, "invalid rank/nranks (max 8 ranks)");
  cudaSetDevice(c->device);
  c->rank = rank;
  c->nranks = nranks;
  if (nranks == 1) return 0;
  if (!g_nccl.load(c->err)) return 1;
  ncclUniqueId_t id;
  memcpy(&id, uniqueId128, 128);
  NK(g_nccl.CommInitRank(&c->comm, nranks, id, rank));
  return 0;
}

int qmc_plan_rebalance(const int *populations, int nRanks, int *sources, int *amounts)
{
  /* Alias-method planner with the semantics of determineLoadBalanceComunicationsAliasMethod
     (dmc/ptools.cpp:93-161): ranks are partitioned into <=k and >k (k = floor(W/P), the first
     r = W mod P positions of the partitioned order get k+1); every under-full rank receives its
     whole deficit from the current over-full rank. */
  if (nRanks < 1) return 1;
  std::vector<int> pop(populations, populations + nRanks), perm(nRanks);
  for (int i = 0; i < nRanks; ++i)
    {
      perm[i] = i;
      sources[i] = 0;
      amounts[i] = 0;
    }
  long total = 0;
  for (int p : pop) total += p;
  const int k = (int)(total / nRanks), r = (int)(total % nRanks);
  int i = 0;
  while (i < nRanks && pop[i] <= k) ++i;
  for (int j = i + 1; j < nRanks; ++j)
    if (pop[j] <= k)
      {
        std::swap(pop[i], pop[j]);
        std::swap(perm[i], perm[j]);
        ++i;
      }
  auto balanced = [&](int idx) { return k + (idx < r ? 1 : 0); };
  int j = 0;
  while (i > j && i < nRanks)
    {
      const int J = perm[j], I = perm[i];
      sources[J] = I;
      amounts[J] = balanced(j) - pop[j];
      pop[i] -= amounts[J];
      if (pop[i] < balanced(i)) ++i;
      ++j;
    }
  return 0;
}

namespace {
// The genealogical records travel with the walker, as the reference's walker messages carry its storage rings
// (walkers.cpp:181-233, ptools.cpp:271-401).  Sender: follow the tail walkers' ancestor links through the event ring and
// pack the records of their lineage; receiver: give every arriving walker a private lineage — slot base+k at EVERY level,
// base = the largest slot count of any live level, linked to itself — so that later hops find exactly those records.
int migrate_records(qmc_ctx *c, const std::vector<int> &src, const std::vector<int> &amt, int n_send, int n_recv, int n_cur)
{
  const int cap = (int)c->cap, DL = c->fw_link_depth, D = c->fw_depth, stride = c->plan.hist_stride;
  const int width = c->sfl_est.empty() ? 0 : 3 * c->plan.nSpecies;
  const long e_last = c->fw_event - 1;
  const int levels = (int)std::min<long>(c->fw_event, DL);
  const int row_levels = c->fw_est.empty() ? 0 : (int)std::min<long>(c->fw_event, D);
  const size_t msz = (size_t)row_levels * stride + (size_t)levels * width;
  const int comp = c->fw_comp_cur;
  const size_t need_msg = (size_t)std::max(n_send, n_recv) * msz * 8, need_chain = (size_t)std::max(n_send, 1) * levels * 4;
  if (need_msg > c->fw_msg_bytes)
    {
      cudaFree(c->fw_msg_send);
      cudaFree(c->fw_msg_recv);
      c->fw_msg_send = c->fw_msg_recv = nullptr;
      if (cudaMalloc(&c->fw_msg_send, need_msg) != cudaSuccess || cudaMalloc(&c->fw_msg_recv, need_msg) != cudaSuccess)
        return fail(c, "rebalance: cannot allocate " + std::to_string(need_msg >> 20) + " MiB for the records of the migrating walkers");
      c->fw_msg_bytes = need_msg;
    }
  if (need_chain > c->fw_chain_bytes)
    {
      cudaFree(c->fw_chain);
      c->fw_chain = nullptr;
      CK(cudaMalloc(&c->fw_chain, need_chain));
      c->fw_chain_bytes = need_chain;
    }
  if (n_send > 0)
    {
      CK(launch_fw_chain(c->stream, c->fw_comp[comp], c->fw_links, DL, cap, e_last, levels, n_cur - n_send, n_send, c->fw_chain));
      CK(launch_fw_records(c->stream, 0, c->fw_rows, D, row_levels, stride, c->sfl_rec, width, c->fw_links, DL, levels, cap, e_last,
                           c->fw_chain, 0, c->fw_comp[comp], 0, n_send, c->fw_msg_send));
      c->launches += 2;
    }
  if (msz > 0)
    {
      NK(g_nccl.GroupStart());
      size_t off = 0;
      for (int r = 0; r < c->nranks; ++r)
        if (amt[r] > 0 && src[r] == c->rank)
          {
            NK(g_nccl.Send(c->fw_msg_send + off, (size_t)amt[r] * msz, NCCL_FLOAT64, r, c->comm, c->stream));
            off += (size_t)amt[r] * msz;
          }
      if (n_recv > 0) NK(g_nccl.Recv(c->fw_msg_recv, (size_t)n_recv * msz, NCCL_FLOAT64, src[c->rank], c->comm, c->stream));
      NK(g_nccl.GroupEnd());
    }
  if (n_recv > 0)
    {
      std::vector<int> slots(DL);
      CK(cudaMemcpyAsync(slots.data(), c->fw_level_slots, (size_t)DL * 4, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      int base = 0;
      for (int l = 0; l < levels; ++l) base = std::max(base, slots[(e_last - l) % DL]);
      if (base + n_recv > cap)
        return fail(c, "rebalance: no free record slots for the arriving walkers' lineages (capacity " + std::to_string(cap) +
                           "): raise the capacity passed to qmc_set_dmc");
      /* arriving walkers are appended at population indices [n_cur - n_send, ...) */
      CK(launch_fw_records(c->stream, 1, c->fw_rows, D, row_levels, stride, c->sfl_rec, width, c->fw_links, DL, levels, cap, e_last, nullptr,
                           base, c->fw_comp[comp], n_cur - n_send, n_recv, c->fw_msg_recv));
      c->launches += 1;
      const int top = base + n_recv;
      for (int l = 0; l < levels; ++l)
        CK(cudaMemcpyAsync(c->fw_level_slots + (e_last - l) % DL, &top, 4, cudaMemcpyHostToDevice, c->stream));
      for (int l = 0; l < row_levels; ++l)      /* the column sums run over every slot of the level */
        CK(cudaMemcpyAsync(c->fw_nslots + (e_last - l) % D, &top, 4, cudaMemcpyHostToDevice, c->stream));
      CK(cudaStreamSynchronize(c->stream));
    }
  return 0;
}
} // namespace

int qmc_set_migration(qmc_ctx *c, int every, double band)
{
  if (!c) return 1;
  if (every < 1 || !(band >= 0.0)) return fail(c, "qmc_set_migration: every >= 1, band >= 0");
  c->mig_every = every;
  c->mig_band = band;
  return 0;
}

int qmc_rebalance(qmc_ctx *c, int64_t *moved)
{
  if (!c) return 1;
  if (moved) *moved = 0;
  if (!c->planned) return fail(c, "qmc_set_dmc must be called first");
  if (c->planning_only) return fail(c, "planning-only context (device -1): no CUDA device, no compute (there is no CPU fallback)");
  if (c->nranks == 1) return 0;
  cudaSetDevice(c->device);
  if (sync_all(c)) return 1;
  Ctl h;
  CK(cudaMemcpyAsync(&h, c->ctl, sizeof h, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  // Fresh populations of every rank, gathered here as the reference does inside isendReceive (ptools.cpp:244-268): the
  // vector carried by the per-step all-reduce is stale after a previous rebalance.  Every rank also publishes its free
  // slots, so that ALL ranks take the same decision (plan or refuse) before any Send/Recv is posted.
  const int nr = c->nranks;
  if (ensure_scratch(c, 0, (size_t)(nr + 1) * 4 * sizeof(int))) return 1;
  int mine[4] = {h.n_cur, h.n_prev, (int)c->cap, 0};
  std::vector<int> all((size_t)nr * 4);
  CK(cudaMemcpyAsync(c->scr_i + (size_t)nr * 4, mine, sizeof mine, cudaMemcpyHostToDevice, c->stream));
  NK(g_nccl.AllGather(c->scr_i + (size_t)nr * 4, c->scr_i, 4, NCCL_INT32, c->comm, c->stream));
  CK(cudaMemcpyAsync(all.data(), c->scr_i, all.size() * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  std::vector<int> pops(nr), src(nr), amt(nr);
  for (int r = 0; r < nr; ++r) pops[r] = all[4 * r];
  qmc_plan_rebalance(pops.data(), nr, src.data(), amt.data());
  for (int r = 0; r < nr; ++r)      /* identical on every rank: nobody enters the exchange if one receiver cannot take its share */
    if (amt[r] > 0 && all[4 * r + 1] + amt[r] > all[4 * r + 2])
      return fail(c, "rebalance: capacity exceeded on rank " + std::to_string(r) + " (" + std::to_string(all[4 * r + 1]) + " slots in use + " +
                         std::to_string(amt[r]) + " arriving > capacity " + std::to_string(all[4 * r + 2]) + ")");
  const size_t dofs = c->plan.dofs, wsz = 2 * dofs + 2;
  int n_send = 0;
  for (int r = 0; r < nr; ++r)
    if (amt[r] > 0 && src[r] == c->rank) n_send += amt[r];
  const int n_recv = amt[c->rank];
  // from here on a local failure would leave the peers blocked in the exchange: abort the communicator so they error out
  auto fatal = [&](const std::string &m) {
    if (c->comm) g_nccl.CommAbort(c->comm);
    c->comm = nullptr;
    return fail(c, m + " (communicator aborted)");
  };
  const size_t need = (size_t)std::max(n_send, n_recv) * wsz * 8;
  if (need > c->slab_bytes)
    {
      cudaFree(c->slab_send);
      cudaFree(c->slab_recv);
      c->slab_send = c->slab_recv = nullptr;
      c->slab_bytes = 0;
      if (cudaMalloc(&c->slab_send, need) != cudaSuccess || cudaMalloc(&c->slab_recv, need) != cudaSuccess)
        return fatal("rebalance: cannot allocate " + std::to_string(need >> 20) + " MiB of migration slabs");
      c->slab_bytes = need;
    }
  int n_cur = h.n_cur;
  // senders ship their TAIL walkers (ptools.cpp:297-303); same kernels as the in-step migration of qmc_step
  if (n_send > 0 && launch_pack_tail(c->stream, c->buf[c->cur], c->parent, c->ctl, (int)dofs, n_send, c->slab_send) != cudaSuccess)
    return fatal("rebalance: pack kernel launch failed");
  NK(g_nccl.GroupStart());
  {
    size_t off = 0;
    for (int r = 0; r < nr; ++r)
      if (amt[r] > 0 && src[r] == c->rank)
        {
          NK(g_nccl.Send(c->slab_send + off, (size_t)amt[r] * wsz, NCCL_FLOAT64, r, c->comm, c->stream));
          off += (size_t)amt[r] * wsz;
        }
    if (n_recv > 0) NK(g_nccl.Recv(c->slab_recv, (size_t)n_recv * wsz, NCCL_FLOAT64, src[c->rank], c->comm, c->stream));
  }
  NK(g_nccl.GroupEnd());
  if (c->genealogy() && c->fw_event > 0 && migrate_records(c, src, amt, n_send, n_recv, n_cur)) return fatal("rebalance: " + c->err);
  n_cur -= n_send;
  if (n_recv > 0)
    { // received walkers land in free physical slots [n_prev, n_prev+n_recv) and are appended to the population
      if (launch_unpack_append(c->stream, c->buf[c->cur], c->parent, c->ctl, (int)dofs, n_send, n_recv, c->slab_recv) != cudaSuccess)
        return fatal("rebalance: unpack kernel launch failed");
      CK(cudaStreamSynchronize(c->stream));
      n_cur += n_recv;
    }
  int np = h.n_prev + n_recv;
  CK(cudaMemcpyAsync(&c->ctl->n_cur, &n_cur, 4, cudaMemcpyHostToDevice, c->stream));
  c->n_host = n_cur;
  CK(cudaMemcpyAsync(&c->ctl->n_prev, &np, 4, cudaMemcpyHostToDevice, c->stream));
  // the whole population vector is refreshed (every rank knows the plan), not only this rank's entry
  std::vector<double> newpops(nr);
  for (int r = 0; r < nr; ++r) newpops[r] = (double)pops[r] + amt[r];
  for (int r = 0; r < nr; ++r)
    if (amt[r] > 0) newpops[src[r]] -= amt[r];
  CK(cudaMemcpyAsync(&c->ctl->gsum[2], newpops.data(), (size_t)nr * 8, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->pops_pending = false;           /* balanced now: the in-step planner waits for the next all-reduce */
  if (n_send > 0 || n_recv > 0)
    {
      c->mig_events += 1;
      c->mig_out += n_send;
      c->mig_in += n_recv;
    }
  c->launches += 2;
  if (moved) *moved = (int64_t)n_send - n_recv;
  return 0;
}

int qmc_get_info(qmc_ctx *c, const char *key, double *value)
{
  if (!c || !key || !value) return 1;
  const std::string k = key;
  if (!c->planned) return fail(c, "qmc_set_dmc must be called first");
  {
    /* plan introspection (host logic tests): "pass<i>.<field>", "onebody<i>.<field>", "triplet<i>.<field>", "est<i>.<field>" */
    int idx = -1;
    char fld[32] = {0};
    const Plan &P = c->plan;
    if (sscanf(key, "pass%d.%31s", &idx, fld) == 2 && idx >= 0 && idx < P.n2b)
      {
        const PairPass &T = P.p2[idx];
        const std::string f = fld;
        int pi = -1;
        if (f == "setA") *value = T.setA;
        else if (f == "setB") *value = T.setB;
        else if (f == "jkind") *value = T.J.kind;
        else if (f == "vkind") *value = T.vkind;
        else if (f == "hist") *value = T.hist;
        else if (f == "v_val") *value = T.v_val;
        else if (f == "v_thr2") *value = T.v_thr2;
        else if (sscanf(fld, "p%d", &pi) == 1 && pi >= 0 && pi < NJP) *value = T.J.p[pi];
        else return fail(c, "unknown info key " + k);
        return 0;
      }
    if (sscanf(key, "onebody%d.%31s", &idx, fld) == 2 && idx >= 0 && idx < P.n1b)
      {
        const OneBodyTerm &T = P.t1[idx];
        const std::string f = fld;
        if (f == "set") *value = T.set;
        else if (f == "jkind") *value = T.J.kind;
        else if (f == "has_pot") *value = T.has_pot;
        else if (f == "half_omega2") *value = T.half_omega2;
        else return fail(c, "unknown info key " + k);
        return 0;
      }
    if (sscanf(key, "triplet%d.%31s", &idx, fld) == 2 && idx >= 0 && idx < P.n3b)
      {
        const TripletPass &T = P.p3[idx];
        const std::string f = fld;
        if (f == "set") *value = T.set;
        else if (f == "jkind") *value = T.J.kind;
        else if (f == "has_pot") *value = T.has_pot;
        else if (f == "hist") *value = T.hist;
        else if (f == "has_constraint") *value = T.has_constraint;
        else if (f == "v_val") *value = T.v_val;
        else return fail(c, "unknown info key " + k);
        return 0;
      }
    if (sscanf(key, "est%d.%31s", &idx, fld) == 2 && idx >= 0 && idx < P.nest)
      {
        const Estimator &T = P.est[idx];
        const std::string f = fld;
        if (f == "kind") *value = T.kind;
        else if (f == "bins") *value = T.bins;
        else if (f == "offset") *value = T.offset;
        else if (f == "norm") *value = T.norm;
        else if (f == "dxinv") *value = T.dxinv;
        else if (f == "q0") *value = T.q0;
        else if (f == "dq") *value = T.dq;
        else if (f == "fx_scale") *value = T.fx_scale;
        else return fail(c, "unknown info key " + k);
        return 0;
      }
  }
  if (c->planning_only && (k == "regs_propagate" || k == "regs_eval" || k == "fp64_peak_tflops" || k == "ctas_per_sm" || k.rfind("time_kernels", 0) == 0 ||
                           k.rfind("propagate_", 0) == 0))
    return fail(c, "planning-only context (device -1): " + k + " needs a device");
  if (k == "dofs") *value = c->plan.dofs;
  else if (k == "n_one_body") *value = c->plan.n1b;
  else if (k == "n_triplet_passes") *value = c->plan.n3b;
  else if (k == "n_estimators") *value = c->plan.nest;
  else if (k == "hist_stride") *value = c->plan.hist_stride;
  else if (k == "global_population")
    { /* pTools::sum(current_walkers.size()) (dmcDriver.cpp:204): the population entry of the per-step all-reduce */
      if (c->planning_only) return fail(c, "planning-only context (device -1): no CUDA device, no compute (there is no CPU fallback)");
      cudaSetDevice(c->device);
      if (c->pipeline == 2 && sync_all(c)) return 1;
      Ctl h;
      CK(cudaMemcpyAsync(&h, c->ctl, sizeof h, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      /* gsum[1] = sum of the ranks' populations after the last step; migration conserves it */
      *value = c->nranks > 1 ? h.gsum[1] : (double)h.n_cur;
    }
  else if (k == "fw_events") *value = (double)c->fw_event;
  else if (k == "fw_depth") *value = c->fw_depth;
  else if (k == "fw_link_depth") *value = c->fw_link_depth;
  else if (k == "fw_ring_bytes") *value = (double)c->fw_depth * (double)c->cap * c->plan.hist_stride * 8.0;
  else if (k == "groups_per_cta") *value = c->cfg.groups_per_cta;
  else if (k == "log_up") *value = c->dmc.log_up;
  else if (k == "log_down") *value = c->dmc.log_down;
  else if (k == "capacity") *value = (double)c->cap;
  else if (k == "grid") *value = c->lean ? c->lean_cfg.grid : c->small ? c->small_grid : c->cfg.grid;
  else if (k == "estimator_grid") *value = c->est_grid;
  else if (k == "estimator_occupancy_target") *value = c->est_occ;
  else if (k == "block") *value = c->cfg.block;
  else if (k == "smem_bytes") *value = (double)(c->lean ? c->lean_cfg.smem : c->cfg.smem);
  else if (k == "warps_per_walker") *value = c->cfg.nwarps_per_group;      /* general kernels (evaluation entry, estimators) */
  else if (k == "pair_kernel_warps_per_walker") *value = c->lean ? c->lean_cfg.nwarps_per_group : 0;
  else if (k == "kernel_variant") *value = c->ksel;
  else if (k == "pipeline") *value = c->pipeline;
  else if (k == "pair_kernel") *value = c->lean ? 1 : 0;
  else if (k == "graph_replays") *value = (double)c->graph_replays;
  else if (k == "small_kernel") *value = c->small ? 1 : 0;
  else if (k == "pair_hist_kernel") *value = c->lean_hist ? 1 : 0;
  else if (k == "small_walkers_per_warp") *value = c->small ? 32 / small_segment_width(c->plan.N[0]) : 0;
  else if (k == "triplet_cyclic") *value = c->trip_cyclic ? 1 : 0;
  else if (k == "triplet_table_entries") *value = c->plan.trip_count;
  else if (k == "pair_kernel_occupancy_target") *value = c->lean_occ;
  else if (k == "migration_events") *value = (double)c->mig_events;
  else if (k == "migrated_out") *value = (double)c->mig_out;
  else if (k == "migrated_in") *value = (double)c->mig_in;
  else if (k == "migrate_every") *value = c->mig_every;
  else if (k == "migrate_band") *value = c->mig_band;
  else if (k == "population_checksum")
    { /* order-independent fingerprint of this rank's population, low 52 bits (exact in a double; additive over ranks mod 2^52) */
      if (c->planning_only) return fail(c, "planning-only context (device -1): no CUDA device, no compute (there is no CPU fallback)");
      cudaSetDevice(c->device);
      if (sync_all(c)) return 1;
      unsigned long long hsum = 0;
      CK(cudaMemsetAsync(c->d_checksum, 0, 8, c->stream));
      CK(launch_population_checksum(c->stream, c->buf[c->cur], c->parent, c->ctl, c->d_checksum));
      CK(cudaMemcpyAsync(&hsum, c->d_checksum, 8, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      *value = (double)(hsum & ((1ull << 52) - 1));
    }
  else if (k == "ctas_per_sm") *value = propagate_occupancy(c->ksel, c->has_hist, c->cfg.smem);
  else if (k == "launches") *value = (double)c->launches;
  else if (k == "sms") *value = c->sms;
  else if (k == "regs_propagate") *value = c->lean ? pair_move_regs(c->lean_vk, c->lean_occ) : c->small ? small_move_regs(c->plan.N[0]) : propagate_regs(c->ksel, false);
  else if (k == "regs_eval") *value = eval_regs(c->ksel, false);
  else if (k == "n_pair_passes") *value = c->plan.n2b;
  else if (k == "time_kernels_on") { c->time_kernels = 1; *value = 1; }
  else if (k == "time_kernels_off") { c->time_kernels = 0; *value = 0; }
  else if (k == "propagate_ms_total") *value = c->prop_ms;
  else if (k == "propagate_launches_timed") *value = (double)c->prop_launches_timed;
  else if (k == "propagate_timers_reset") { c->prop_ms = 0; c->prop_launches_timed = 0; *value = 0; }
  else if (k == "fp64_peak_tflops")
    {
      cudaSetDevice(c->device);
      *value = measure_fp64_peak_tflops(c->stream, nullptr);
    }
  else return fail(c, "unknown info key " + k);
  return 0;
}

void *qmc_stream(qmc_ctx *c) { return c ? (void *)c->stream : nullptr; }

} // extern "C"
