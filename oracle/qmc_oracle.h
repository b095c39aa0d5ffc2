/* qmc_oracle.h — plain-C restatement of the reference's DMC hot path.  TEST INFRASTRUCTURE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * link or load this; the product never does.  Pinned against the reference itself
 * (oracle/_ref, built from the unmodified sources) and against the golden vectors generated
 * through it (tests/golden/) by tests/test_oracle.py.
 *
 * The model is a flat term list (the Python loader oracle/oracle.py flattens the reference's
 * input.json into it, using the reference's id strings):
 *   ti[8*t + 0] = term class   1 wavefunction, 2 potential, 3 estimator
 *   ti[8*t + 1] = kind         (ORC_WF_*, ORC_POT_*, ORC_EST_*)
 *   ti[8*t + 2..4] = sets
 *   ti[8*t + 5] = radial functor kind (ORC_J_*) for wavefunction terms; bins for estimators
 *   ti[8*t + 6] = number of parameters,  ti[8*t + 7] = offset of the parameters in tp[]
 * Wavefunction parameters are the JSON-level numbers of the reference constructors (same order
 * as include/qmc4_b200.h documents), followed by hardSphereRadius (or -1) for 3-body terms.
 * Scalar observables are 1-bin estimators: sums[0] = sum of the observable, weight = walkers (accumulators.h:21-47).
 */
#ifndef QMC_ORACLE_H
#define QMC_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_J_GAUSSIAN = 1, ORC_J_GAUSSGAUSS, ORC_J_SQUAREWELL, ORC_J_HARDSPHEREGAUSS, ORC_J_BSPLINE,
       ORC_J_POSCHTELLER, ORC_J_HARDSPHERE, ORC_J_SQUAREWELL_BOUND,
       ORC_J_DELTA_PHONONS, ORC_J_DELTA_IN_TRAP, ORC_J_DELTA_BOUND_STATE_PHONONS, ORC_J_DIPOLAR_REP };
enum { ORC_WF_JASTROW1B = 1, ORC_WF_JASTROW2B_UNDIS, ORC_WF_JASTROW2B_DIS, ORC_WF_JASTROW3B_UNDIS, ORC_WF_JASTROW3B_DIS };
enum { ORC_POT_HARMONIC1B = 1, ORC_POT_SQUAREWELL2B, ORC_POT_DIPOLAR2B, ORC_POT_POSCHTELLER2B,
       ORC_POT_TWOBODY_SQUAREWELL, ORC_POT_TWOBODY_BARRIER, ORC_POT_THREEBODY_UNDIS_SQUAREWELL,
       ORC_POT_THREEBODY_UNDIS_BARRIER, ORC_POT_THREEBODY_DIS_SQUAREWELL, ORC_POT_THREEBODY_DIS_BARRIER };
enum { ORC_EST_PAIRCORRELATION = 1, ORC_EST_TRIMERCORRELATION_UNDIS, ORC_EST_STRUCTUREFACTOR, ORC_EST_TRIMERCORRELATION,
       ORC_EST_TRIMERCORRELATION_DIS, ORC_EST_TRIMERHYPERRADIUS /* scalar: bins = 1 */, ORC_EST_CENTEROFMASSSQUARED /* scalar */ };

typedef struct orc_model {
  double lBox[3];
  int nSpecies;
  const int *N;
  int nTerms;
  const int *ti;
  const double *tp;
} orc_model;

const char *orc_last_error(void);

int orc_eval_model(const orc_model *m, const double *R, int nW, double *logpsi, double *grad, double *lap,
                   double *V, double *E, int *valid);
int orc_green_ratio_model(const orc_model *m, double timeStep, int nW, const double *Rnew, const double *Fnew,
                          const double *Rold, const double *Fold, double *q);
int orc_drift_diffuse_model(const orc_model *m, double timeStep, int nW, const double *Rold, const double *Fold,
                            const double *xi_draw_order, double *Rnew);
int orc_histogram_model(const orc_model *m, const double *R, int nW, int iEst, double *sums, int bins, double *weight);
int orc_forward_walking_model(const orc_model *m, const double *R, const int *n, const int *anc, int K, int iEst,
                              int recordSteps, int fwSteps, double *sums, int bins, double *weight);
double orc_bessel_k(int n, double x);
void orc_bessel_k01e(double x, double *k0, double *k1);
int orc_force_energy_model(const orc_model *m, const double *R, int nW, double *out);
int orc_superfluid_fraction_model(const orc_model *m, const double *R, const int *n, const int *anc, int K,
                                  const int *storedSets, int nStored, int recordSteps, int idxA, int idxB,
                                  double *sums, double *weights, int bins);

/* branchingControl (dmc/branching.cpp) with the uniforms supplied by the caller */
double orc_energy_shift(const double *E, int n, double timeStep, double meanWalkers, double deltaWalkers);
int orc_n_descendants(double Enew, double Eold, double energyShift, double timeStep, double u);
int orc_branch(const int *nDesc, const double *E, int n, double *Eafter, int capOut);
int orc_plan(const int *populations, int nRanks, int *sources, int *amounts);

/* one full dmcDriver::update (dmc/dmcDriver.cpp:56-84) per walker given the random numbers:
 * xi in draw order, u_m for Metropolis.  Outputs the post-move walker (R,F,logpsi,E) and flags. */
int orc_update_model(const orc_model *m, double timeStep, int nW, const double *Rold, const double *xi_draw_order,
                     const double *u_m, double *Rout, double *Fout, double *logpsi_out, double *E_out, int *accepted);

/* jastrow<T>::print(minx,maxx,n) (jastrow.h:42-63): n rows {x, u, u', u''}; p = JSON-level parameters of the functor */
int orc_tabulate(int jastrowKind, const double *p, int np, double minx, double maxx, int n, double *out);

/* one vmcDriver::step (dmc/vmcDriver.cpp:55-84) per chain given xi (draw order) and the Metropolis uniform */
int orc_vmc_step_model(const orc_model *m, double sigma, int nW, const double *Rold, const double *xi_draw_order,
                       const double *u_m, double *Rout, double *logpsi_out, double *E_out, int *accepted);

#ifdef __cplusplus
}
#endif
#endif
