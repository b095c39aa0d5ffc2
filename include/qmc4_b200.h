/* qmc4_b200.h — C ABI of the B200-native DMC walker-propagation hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.
 * A host driver that keeps the reference's JSON input and factory API (qmc4_b200/host/) — or
 * the reference itself, see INTEGRATION.md — turns every object its factories create into one
 * qmc_add_* call, then drives qmc_step()/qmc_block_results() where the reference runs
 * dmcDriver::step()/out().
 *
 * Every entry point cites the reference interface it replaces (paths under /root/reference).
 * All functions return 0 on success; on failure qmc_last_error() describes the problem (the
 * C++ host wrapper rethrows it as the reference's exception types, dmc/qmcExceptions.h).
 * The library is re-entrant per context; a context is bound to one CUDA device and must be
 * used from one host thread at a time.  There is NO CPU fallback: without a usable CUDA
 * device qmc_create() fails.
 *
 * Walker layout (identical to the reference's column-major Eigen state_t, dmc/traits.h:15, and
 * to what MPI ships per walker, dmc/walkers.cpp:89-192): one walker = for each species
 * s = 0..S-1, contiguously, data[d*N_s + i] (d = 0..2).  `dofs` = 3*sum(N_s) doubles.
 */
#ifndef QMC4_B200_H
#define QMC4_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QMC4_B200_ABI_VERSION 1

typedef struct qmc_ctx qmc_ctx;

/* ---- radial Jastrow functors u(r) (dmc/wavefunction/jastrows/) -------------------------------
 * params are the JSON-level numbers of the reference constructors, in the order listed;
 * derived coefficients are computed by the library exactly as the reference constructors do. */
enum qmc_jastrow_kind {
  QMC_J_GAUSSIAN = 1,        /* "gaussian"        {alpha}                        jastrowGaussian.h:23-25       */
  QMC_J_GAUSSGAUSS = 2,      /* "gaussiangaussian"{r0,u0}                        jastrowGaussGauss.h:19-21     */
  QMC_J_SQUAREWELL = 3,      /* "squareWell"      {V0,R0,Rm,alpha,aInverse,cut_off}  jastrowSquareWell.h:20-62 */
  QMC_J_HARDSPHEREGAUSS = 4, /* "hardSphereGauss" {a,Rm,D}                       jastrowHardSphereGauss.h:13-23*/
  QMC_J_BSPLINE = 5,         /* "bSpline"         {stepSize,valueRight,derivativeRight,c_0..c_{n-1}} jastrowSpline.cpp:80-110 */
  QMC_J_POSCHTELLER = 6,     /* "poschTeller"     {R0,C,alpha,Rm,cut_off}        jastrowPoschTeller.h:19-57    */
  QMC_J_HARDSPHERE = 7,      /* "hardSphere"      {a,k,cut_off}                  jastrowHardSphereGauss.h:39-60*/
  QMC_J_SQUAREWELL_BOUND = 8,/* "squareWellBoundState" {V0,R0,Rm,cut_off,aInverse,eb,alpha} jastrowSquareWell.h:80-150 */
  QMC_J_DELTA_PHONONS = 9,   /* "delta_phonons"   {k,delta,beta,z,cut_off}       jastrow_delta.h:4-54 (g is parsed, unused: process() is never called) */
  QMC_J_DELTA_IN_TRAP = 10,  /* "delta_in_trap"   {a}                            jastrow_delta.h:56-72         */
  QMC_J_DELTA_BOUND_STATE_PHONONS = 11,/* "delta_bound_state_phonons" {xI,k,A,beta,cut_off} jastrow_delta.h:74-98 */
  QMC_J_DIPOLAR_REP = 12     /* "dipolar_rep"     {D,cut_off,matching_point,alpha,C}  jastrowDipolar.h:9-73 (Bessel K_n on the
                                device instead of boost::math::cyl_bessel_k; see DESIGN.md for how it is pinned) */
};

/* ---- wavefunction terms (dmc/wavefunction/) ------------------------------------------------- */
enum qmc_wavefunction_kind {
  QMC_WF_JASTROW1B = 1,       /* "jastrow1b/<J>"      sets={A}      jastrowWavefunctionOneBody.h:53-85   */
  QMC_WF_JASTROW2B_UNDIS = 2, /* "jastrow2bUnDis/<J>" sets={A,A}    jastrowWavefunctionTwoBody.h:124-164 */
  QMC_WF_JASTROW2B_DIS = 3,   /* "jastrow2bDis/<J>"   sets={A,B}    jastrowWavefunctionTwoBody.h:33-76   */
  QMC_WF_JASTROW3B_UNDIS = 4, /* "jastrow3bUnDis/<J>" sets={A}      jastrowWavefunctionThreeBody.h:91-137*/
  QMC_WF_JASTROW3B_DIS = 5    /* "jastrow3bDis/<J>"   sets={A,B,C} (all different)  jastrowWavefunctionThreeBody.h:164-305;
                                 hardSphereRadius -> hardSphereConstraintThreeBodyDis (jastrowWavefunctionThreeBody.cpp:39-80) */
};

/* ---- potentials (dmc/potential.cpp, potentialTwoBody.h, potentialThreeBody.h) ---------------- */
enum qmc_potential_kind {
  QMC_POT_HARMONIC1B = 1,        /* "harmonic1b"   sets={A}    params={omega}     potential.cpp:21-32   */
  QMC_POT_SQUAREWELL2B = 2,      /* "squareWell2b" sets={A,B}  params={V0,R0}     potential.cpp:57-65   */
  QMC_POT_DIPOLAR2B = 3,         /* "dipolar2b"    sets={A,B}  params={D}         potential.cpp:72-98   */
  QMC_POT_POSCHTELLER2B = 4,     /* "poschTeller2b" sets={A,B} params={R0}        potential.cpp:100-114 */
  QMC_POT_TWOBODY_SQUAREWELL = 5,  /* "potentialTwoBody/squareWell"        sets={A,B} params={R0,V0}  potentialTwoBody.h:23-33, potentialFunctors.h:5-19  */
  QMC_POT_TWOBODY_BARRIER = 6,     /* "potentialTwoBody/potentialBarrier"  sets={A,B} params={R0,V0}  potentialFunctors.h:22-36 */
  QMC_POT_THREEBODY_UNDIS_SQUAREWELL = 7, /* "potentialThreeBodyUnDis/squareWell"       sets={A,A,A} params={R0,V0} potentialThreeBody.h:23-44 */
  QMC_POT_THREEBODY_UNDIS_BARRIER = 8,    /* "potentialThreeBodyUnDis/potentialBarrier" sets={A,A,A} params={R0,V0} */
  QMC_POT_THREEBODY_DIS_SQUAREWELL = 9,   /* "potentialThreeBodyDis/squareWell"         sets={A,B,C} params={R0,V0} potentialThreeBody.h:60-125 */
  QMC_POT_THREEBODY_DIS_BARRIER = 10      /* "potentialThreeBodyDis/potentialBarrier"   sets={A,B,C} params={R0,V0} */
};

/* ---- estimators accumulated every sampling step (dmc/estimatorsFactory.cpp:39-133) ----------- */
enum qmc_estimator_kind {
  QMC_EST_PAIRCORRELATION = 1,       /* "pairCorrelation"        sets={A,B}  pairCorrelation.cpp:17-59   */
  QMC_EST_TRIMERCORRELATION_UNDIS = 2, /* "trimerCorrelationUnDis" sets={A}  trimerCorrelation.cpp:82-143 */
  QMC_EST_STRUCTUREFACTOR = 3,       /* "structureFactor" sets={A,B} extra={maxq,lBox,channel}; bins q-points
                                        (structureFactor.cpp:8-90, applied to coordinate 0 — the reference
                                        compiles it for DIMENSIONS==1 only) */
  QMC_EST_TRIMERCORRELATION = 4,     /* "trimerCorrelation"    sets={A,B,C}, N_C must be 1 (assert, trimerCorrelation.cpp:11);
                                        normalisation 1/(dx 4 pi N_A N_B N_B) as in the reference (:69, sic) */
  QMC_EST_TRIMERCORRELATION_DIS = 5, /* "trimerCorrelationDis" sets={A,B,C}  trimerCorrelation.cpp:163-227 */
  /* scalar observables (realScalarEstimator, estimators.cpp:35-48,89-98): bins must be 1, minx/maxx unused;
     qmc_get_estimator returns sums[0] = sum of the observable over the accumulated walkers and their number in
     *weight (scalarAccumulator, accumulators.h:21-47: average = sum/n) */
  QMC_EST_TRIMERHYPERRADIUS = 6,     /* "trimerhyperRadius"    sets={A,B,C}, N_C must be 1   hyperRadius.cpp:17-43 */
  QMC_EST_CENTEROFMASSSQUARED = 7    /* "centerOfMassSquared"  sets={A}   centerOfMassSquared.cpp:8-18: (sum_i |r_i|)^2/N */
};

/* Results of one block, the numbers dmcDriver::out() prints and dumps (dmc/dmcDriver.cpp:197-230). */
typedef struct qmc_block_result {
  double energy_sum;      /* sum of E_L over walkers x sampling steps (scalarAccumulator, accumulators.h:27) */
  double energy_count;    /* number of samples                                                          */
  double n_moves;         /* metropolis::n          (metropolis.cpp:26)                                 */
  double n_accepted;      /* metropolis::nAccepted  (metropolis.cpp:24)                                 */
  double population;      /* current walkers on this rank                                               */
  double energy_shift;    /* E_T currently in use (branching.cpp:65-97)                                 */
  double steps_done;      /* DMC steps executed since the previous call                                 */
  double overflow;        /* !=0: population hit the buffer capacity and copies were dropped            */
  double device_ms;       /* CUDA-event time of the steps since the previous call                       */
} qmc_block_result;

/* Per-walker trace of one propagate step, for parity tests (all device->host copies). */
typedef struct qmc_trace {
  double *xi;        /* [nW*dofs] unit normal deviates, walker layout                         */
  double *Rprop;     /* [nW*dofs] proposed positions R' (dmcMoves.cpp:12-30)                  */
  double *Fprop;     /* [nW*dofs] grad log psi at R'                                          */
  double *logpsi;    /* [nW] log psi(R')                                                      */
  double *lap;       /* [nW]                                                                   */
  double *V;         /* [nW]                                                                   */
  double *Eprop;     /* [nW] E_L(R')                                                          */
  double *q;         /* [nW] log Green ratio (dmcMoves.cpp:32-57)                             */
  double *u_metropolis; /* [nW] uniform used by acceptLog (metropolis.cpp:15)                 */
  double *u_branch;  /* [nW] uniform used by nDescendants (branching.cpp:13)                  */
  int    *accepted;  /* [nW]                                                                   */
  int    *valid;     /* [nW] constraints satisfied (productWavefunction.cpp:70-82)            */
  int    *n_desc;    /* [nW]                                                                   */
  int    *parent;    /* [nW] source slot in the pre-step population                           */
} qmc_trace;

const char *qmc_last_error(const qmc_ctx *ctx);   /* ctx may be NULL: error of a failed qmc_create */
int  qmc_abi_version(void);

/* geometryPBC(lBox) + particle numbers N[] (dmc/main.cpp:160-180). */
int  qmc_create(qmc_ctx **ctx, int device, const double lBox[3], int nSpecies, const int *N);
void qmc_destroy(qmc_ctx *ctx);

/* One call per object created by wavefunctionFactory::create (wavefunctionFactory.h:35-49).
 * hardSphereRadius < 0: none ("hardSphereRadius" key, jastrowWavefunctionThreeBody.h:52-61). */
int  qmc_add_wavefunction_term(qmc_ctx *ctx, int wf_kind, const int *sets, int nsets,
                               int jastrow_kind, const double *params, int nparams,
                               double hardSphereRadius);
/* One call per object created by potentialFactory::create (potentialFactory.h:43-55). */
int  qmc_add_potential_term(qmc_ctx *ctx, int pot_kind, const int *sets, int nsets,
                            const double *params, int nparams);
/* One call per estimator created by estimatorFactory::create; returns its index in *est_index.
 * (the energy estimator is implicit, dmcDriver.cpp:245). */
int  qmc_add_estimator(qmc_ctx *ctx, int est_kind, const int *sets, int nsets,
                       int bins, double minx, double maxx,
                       const double *extra, int nextra, int *est_index);

/* Forward walking (pure estimators).  Replaces, for the histogram estimator `est_index`:
 *   storerFactory::create (estimatorsFactory.h:97-131, "recordSteps")  -> realHistogramStorer
 *     (correlationEstimator.cpp:81-125: the walker's own histogram stored every sampling step in a
 *      ring of recordSteps entries that follows the walker through branching, walkers.cpp:139-179),
 *   estimatorFactory::create (estimatorsFactory.cpp:47-74, "forwardWalkingSteps": [..])
 *     -> one realHistogramForwardWalkingEstimator per entry (correlationEstimator.cpp:128-175): adds,
 *      for every walker of the current population, the record its ancestor stored fw sampling
 *      steps earlier, once recordSteps records exist.
 * The n_fw estimators get the indices *first_index .. *first_index+n_fw-1 (>= QMC_FW_INDEX_BASE) for
 * qmc_get_estimator.  Must precede qmc_set_dmc.  Errors: vector observables (structure factor) have
 * no forward-walking estimator in the reference either (missingImplementation, :68);
 * fw >= recordSteps (assert, correlationEstimator.cpp:137).  With a communicator qmc_rebalance ships the lineage records
 * of the migrating walkers with them (the reference's walker messages carry the rings, walkers.cpp:181-233) and
 * qmc_get_estimator returns the sums reduced over the ranks. */
#define QMC_FW_INDEX_BASE 65536
int  qmc_add_forward_walking(qmc_ctx *ctx, int est_index, int record_steps,
                             const int *fw_steps, int n_fw, int *first_index);

/* "forceEnergy" measurement (main.cpp:270-273: realScalarEstimator("forceEnergy", new forceEnergy(&pot)), attached to
 * vmcDriver (:318-335) and to dmcDriver for svmc only (:395-408)).  forceEnergy::operator() (energy.cpp:37-61) =
 * 0.5 * sum_i |grad_i log psi|^2 + V of every walker at every sampling step, weight 1.  Device: one extra evaluation
 * launch + reduction per sampling step, outside the hot kernels.  Must precede qmc_set_dmc(branching_on = 0) /
 * qmc_set_vmc; qmc_set_dmc with branching fails, as the reference never attaches it there.
 * qmc_get_force_energy: {sum, count} since the last qmc_clear_estimators (reduced over ranks). */
int  qmc_enable_force_energy(qmc_ctx *ctx);
int  qmc_get_force_energy(qmc_ctx *ctx, double *sum, double *count);

/* "superfluidFraction" measurement (estimatorsFactory.cpp:80-116).  Replaces
 *   createSuperfluidFractionStorer + centerOfMassStorer (correlationEstimator.cpp:187-270,363-410): every walker
 *     carries a ring of `bins` records {sum_i x_i^d of each stored set}, written at every sampling step, copied
 *     with the walker by branching (here: one record per surviving slot, found through the ancestor links
 *     forward walking uses — nothing is copied);
 *   superfluidFractionEstimator::accumulate (:322-357): for the record j written last and every record i <= j of
 *     the same turn of the ring, sums[j-i] += sum_d (XA_i - XA_j)(XB_i - XB_j), weights[j-i] += 1, per walker.
 * setA/setB are species indices (the library stores all species; the factory's remapping through the storer's
 * set list, :86-108, is not needed).  All measurements must use the same `bins` (the reference derives the record
 * stride from it, :333).  Index >= QMC_SFL_INDEX_BASE for qmc_get_vector_estimator.  Must precede qmc_set_dmc; not
 * with qmc_set_vmc (main.cpp:279-282).  Multi-rank: records migrate in qmc_rebalance, results are reduced over ranks. */
#define QMC_SFL_INDEX_BASE 131072
int  qmc_add_superfluid_fraction(qmc_ctx *ctx, int setA, int setB, int bins, int *est_index);
/* realVectorAccumulator state of a superfluidFraction estimator since the last qmc_clear_estimators: sums[bins],
 * weights[bins]; the reference writes "i sums[i]/weights[i]" for weights[i] > 0 (correlationEstimator.cpp:305-320). */
int  qmc_get_vector_estimator(qmc_ctx *ctx, int est_index, double *sums, double *weights, int bins);

/* dmcDriver constructor arguments + seed + performBranching (dmcDriver.cpp:232-255,
 * main.cpp:363-395).  capacity = walker slots to allocate on this rank (0: 2*(target+delta)/nranks+64).
 * Builds the kernel plan: must follow all qmc_add_* calls. */
int  qmc_set_dmc(qmc_ctx *ctx, double timeStep, int64_t targetWalkers, int64_t deltaWalkers,
                 uint64_t seed, int branching_on, int64_t capacity);

/* Many-chain VMC in place of qmc_set_dmc: every walker is an independent Metropolis chain of vmcDriver
 * (vmcDriver.cpp:12-84): gaussianMover x' = x + xi*sqrt(sigma) with sigma = the input's "timeStep"
 * (vmcMoves.cpp:11-35, main.cpp:314), acceptance min(1, |psi(x')/psi(x)|^2) (vmcDriver.cpp:68,
 * metropolis.cpp:5-29), constraint violation = rejection (:73-77), no branching; the energy estimator
 * is the local energy of the current configuration (energy.cpp:28-35).  The reference runs ONE chain
 * per MPI rank; nChains > 1 is the GPU extension (statistically equivalent: chains are independent). */
int  qmc_set_vmc(qmc_ctx *ctx, double sigma, int64_t nChains, uint64_t seed, int64_t capacity);

/* wavefunction::print() -> jastrow<T>::print(minx, maxx, n) (wavefunction.h:64, jastrow.h:42-63; main.cpp:220-227
 * writes it to wave<i>.dat): u, u', u'' of the radial functor of wavefunction term `wf_index` (order of the
 * qmc_add_wavefunction_term calls) evaluated ON THE DEVICE at the n bin centres x_i = minx + (i+0.5)(maxx-minx)/n.
 * out (host): n rows {x, u, u', u''}. */
int  qmc_tabulate_wavefunction(qmc_ctx *ctx, int wf_index, double minx, double maxx, int64_t n, double *out);

/* Parity entry (1): evaluate nW configurations given in HOST memory; any output may be NULL.
 * = initializer::initialize + updateForceGradientEnergy (initializer.cpp:75-82, walkers.cpp:46-61). */
int  qmc_eval(qmc_ctx *ctx, const double *R, int64_t nW,
              double *logpsi, double *grad, double *lap, double *V, double *E, int *valid);
/* Same on DEVICE pointers, asynchronous on the context stream (bench: inputs resident in HBM). */
int  qmc_eval_device(qmc_ctx *ctx, const double *dR, int64_t nW,
                     double *dlogpsi, double *dgrad, double *dlap, double *dV, double *dE, int *dvalid);
/* Estimators over nW given configurations with unit weight each (= dmcDriver::accumulate over
 * them, dmcDriver.cpp:257-264); results are read with qmc_get_estimator. */
int  qmc_accumulate_configs(qmc_ctx *ctx, const double *R, int64_t nW);

/* Population: host->device upload of nW walkers (dmcDriver::run's `states`, dmcDriver.cpp:160-173),
 * then initial evaluation + first E_T (dmcDriver.cpp:163,191).  Fails with "Invalid initial
 * configurations" like the reference if a constraint is violated. */
int  qmc_upload_walkers(qmc_ctx *ctx, const double *R, int64_t nW);
int  qmc_download_walkers(qmc_ctx *ctx, double *R, int64_t capacityWalkers, int64_t *nW,
                          double *E /* may be NULL */);

/* nSteps x { dmcDriver::step(); if ((k+1) % sampleEvery == 0) accumulate(); } enqueued
 * asynchronously (driver.cpp:28-39: sampleEvery = correlationSteps+1). */
int  qmc_step(qmc_ctx *ctx, int nSteps, int sampleEvery);
/* One step with a per-walker trace (host buffers, any may be NULL); synchronous. */
int  qmc_step_traced(qmc_ctx *ctx, int accumulate, const qmc_trace *trace, int64_t traceCapacity,
                     int64_t *nW_before, int64_t *nW_after);
/* dmcDriver::out(): synchronise, return and clear the block accumulators (dmcDriver.cpp:197-230).
 * With a communicator the sums are global on every rank (the reference reduces to rank 0). */
int  qmc_block_results(qmc_ctx *ctx, qmc_block_result *out);
/* Histogram/vector estimator state since the last qmc_clear_estimators: sums[bins], weight
 * (accumulators.h:101-160); with a communicator reduced over ranks. */
int  qmc_get_estimator(qmc_ctx *ctx, int est_index, double *sums, int bins, double *weight);
int  qmc_clear_estimators(qmc_ctx *ctx);
int  qmc_synchronize(qmc_ctx *ctx);

/* Multi-GPU (replaces dmc/ptools.cpp): one context per GPU/rank.  uniqueId = the 128-byte
 * ncclUniqueId created by rank 0 and distributed by the launcher (torch.distributed / MPI). */
int  qmc_comm_unique_id(void *uniqueId128);
int  qmc_comm_init(qmc_ctx *ctx, const void *uniqueId128, int rank, int nranks);
/* walkerDistribution::isendReceive + wait (ptools.cpp:271-401): allgather populations, alias-method
 * plan, ship tail walkers to under-full ranks.  moved = walkers sent (+) / received (-).
 * With forward walking / superfluidFraction the walkers' records travel too (the reference's walker messages carry the
 * storage rings, walkers.cpp:181-233): per walker (max fw + 1) x bins + bins_sfl x 3 nSpecies doubles; the receiver needs
 * free record slots above the largest population of the last `depth` sampling events (error if the capacity is too small). */
int  qmc_rebalance(qmc_ctx *ctx, int64_t *moved);
/* Inside qmc_step the same migration runs by itself, ONE STEP LATE and overlapped with the next step's moves exactly as
 * dmcDriver::step does (isend after step k, dmcDriver.cpp:112,150-157; wait after the local loop of step k+1, arrivals
 * advanced within that step, :133-136): the plan is derived on every rank from the population vector the per-step
 * all-reduce carries, tail walkers travel on a second stream while the move kernel of the next step runs.
 * every: plan every `every` steps (default 1, like the reference).  band: dead band — transfers are posted only when
 * max-min population exceeds band x mean population (default 0.02; 0 = any imbalance, like the reference). */
int  qmc_set_migration(qmc_ctx *ctx, int every, double band);
/* The planner alone (ptools.cpp:93-161), host-side, for tests. */
int  qmc_plan_rebalance(const int *populations, int nRanks, int *sources, int *amounts);

/* Introspection used by bench.py / tests. */
int  qmc_get_info(qmc_ctx *ctx, const char *key, double *value);
void *qmc_stream(qmc_ctx *ctx);     /* cudaStream_t the context launches on */

#ifdef __cplusplus
}
#endif
#endif /* QMC4_B200_H */
