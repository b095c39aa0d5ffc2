// plan.h — the POD "kernel plan" the term-list compiler produces from the objects the reference's
// factories would create (SURVEY.md §7 stage 1).  It replaces per-walker virtual dispatch
// (dmc/wavefunction/wavefunction.h:34-81, dmc/potential.h:12-33) by one flat description that is
// passed to every kernel as a __grid_constant__ parameter (constant bank, uniform loads).
#pragma once
#include <stdint.h>

namespace qmc {

constexpr int MAX_SPECIES = 4;
constexpr int MAX_ONEBODY = 4;
constexpr int MAX_PAIRPASS = 4;
constexpr int MAX_TRIPLETPASS = 3;
constexpr int MAX_EST = 6;
constexpr int NJP = 16;   // derived parameters per radial functor

// radial functor kinds as evaluated on the device (qmc_jastrow_kind + NONE)
enum : int { J_NONE = 0, J_GAUSSIAN = 1, J_GAUSSGAUSS = 2, J_SQUAREWELL = 3, J_HARDSPHEREGAUSS = 4,
             J_BSPLINE = 5, J_POSCHTELLER = 6, J_HARDSPHERE = 7, J_SQUAREWELL_BOUND = 8,
             J_DELTA_PHONONS = 9, J_DELTA_IN_TRAP = 10, J_DELTA_BOUND_STATE_PHONONS = 11, J_DIPOLAR_REP = 12,
             J_GENERIC = 99 /* template tag: switch on the runtime kind */ };
// pair potential kinds on the device
enum : int { V_NONE = 0, V_STEP = 1 /* r<=R0 ? v : 0 (squareWell: v=-V0, barrier: v=+V0) */,
             V_DIPOLAR = 2, V_POSCHTELLER = 3, V_GENERIC = 99 };

struct RadialFn {          // one radial Jastrow functor with the reference's derived coefficients
  int kind;
  int ncoef;               // bSpline: number of coefficients (in Plan::spline at `coef_off`)
  int coef_off;
  int pad;
  double p[NJP];
};

struct OneBodyTerm {       // jastrow1b (jastrowWavefunctionOneBody.h:53-85) and/or harmonic1b (potential.cpp:21-32)
  int set;
  int has_pot;             // harmonic: V += 0.5*omega^2*r^2
  double half_omega2;
  RadialFn J;              // kind J_NONE if only the potential is present
};

struct PairPass {          // one fused pass over the (setA,setB) pair table
  int setA, setB;          // setA==setB: triangular (UnDis)   else rectangular (Dis)
  int vkind;
  int hist;                // index into Plan::est of a pairCorrelation on the same sets, or -1
  double v_thr2;           // V_STEP: largest r2 with sqrt(r2) <= R0 (bit-exact threshold)
  double v_val;            // V_STEP: value inside; V_DIPOLAR: D; V_POSCHTELLER: alpha
  RadialFn J;              // J_NONE if the pass only carries a potential / estimator
};

struct TripletPass {       // jastrow3bUnDis + potentialThreeBodyUnDis + trimerCorrelationUnDis on set A, or their
                           // distinguishable versions (jastrow3bDis, potentialThreeBodyDis, trimerCorrelation{,Dis}) on (A,B,C)
  int set;                 // A
  int has_pot;
  int hist;                // index into Plan::est or -1
  int has_constraint;
  int setB, setC;          // -1: undistinguishable pass on `set`; else three different sets (LOOP3B_DIS, tools3B.h:25-43)
  double v_R0, v_val;      // F(R) = R<=R0 ? v_val : 0
  double hs_R0;            // hardSphereConstraintThreeBodyUnDis: invalid if any R <= hs_R0
  double v_thr2, hs_thr2;  // largest R^2 whose correctly rounded sqrt is <= v_R0 / hs_R0 (bit-exact thresholds on R^2)
  RadialFn J;              // may be J_NONE
};

struct Estimator {
  int kind;                // qmc_estimator_kind
  int setA, setB;
  int bins;
  int offset;              // offset (doubles) of this estimator's sums inside one histogram replica
  int channel;             // structureFactor
  int setC;                // three-set observables (trimerCorrelation, trimerCorrelationDis, trimerhyperRadius)
  int pad;
  double minx, maxx, dxinv;
  double norm;             // pairCorrelation.cpp:41-59 / trimerCorrelation.cpp:128-143 normalisation factor
  double q0, dq;           // structureFactor q-grid (structureFactor.cpp:8-40)
  double fx_scale;         // pairCorrelation: fixed-point scale of 1/r^2 (2^36 * maxx^2), 0 = fp64 atomics only
};

struct Plan {
  int nSpecies;
  int Ntot;                // particles per walker (all species)
  int dofs;                // 3*Ntot
  int N[MAX_SPECIES];
  int off[MAX_SPECIES];    // offset (doubles) of species s inside a walker = 3*sum_{t<s} N_t
  double L[3], Linv[3];
  int n1b, n2b, n3b, nest;
  OneBodyTerm t1[MAX_ONEBODY];
  PairPass p2[MAX_PAIRPASS];
  TripletPass p3[MAX_TRIPLETPASS];
  Estimator est[MAX_EST];
  int hist_stride;         // doubles per histogram replica
  int hist_replicas;
  int needs_pair_table;    // doubles of shared memory reserved per group for the triplet pair tables (0: none)
  int aux_doubles;         // doubles per group for structure-factor partial sums (nwarps*2*bins), set at launch config
  // cyclic triplet schedule (run_triplet_pass_cyclic, evaluate.cuh): canonical (p, q) offsets modulo trip_mod, every unordered
  // triplet of a set of trip_mod <= 32 particles exactly once; device pointer, entry = p | q << 8 | symmetric << 16
  const int *trip_tab;
  int trip_count, trip_mod;
};

// DMC run parameters (dmcDriver.cpp:232-255, branching.cpp:5-8)
struct DmcParams {
  double tau, sqrt_tau;
  double mean_walkers, delta_walkers;
  double log_up, log_down;   // (1/tau)*log(N0/(N0+D)), (1/tau)*log(N0/(N0-D))  (branching.cpp:84-96)
  uint32_t seed_lo, seed_hi;
  int branching_on;
  int rank, nranks;
  int vmc;                   // 1: many-chain VMC (vmcDriver.cpp:55-84): Gaussian move without drift (tau = 0, sqrt_tau = sqrt(sigma),
                             //    vmcMoves.cpp:11-35), acceptance |psi'|^2/|psi|^2 only (no Green-function ratio), no branching
};

} // namespace qmc
