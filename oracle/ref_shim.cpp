/* oracle/ref_shim.cpp — C entry points over the UNMODIFIED reference classes.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is compiled together with the reference's own
 * /root/reference/dmc/ *.cpp (in place, never copied) into oracle/_ref/libqmc4ref.so by
 * oracle/Makefile.  It is the "reference itself run here" arm of the oracle: the plain-C
 * restatement in oracle/qmc_oracle.c and the CUDA product are both checked against it, and the
 * golden vectors under tests/golden/ are generated through it (tests/golden/make_golden.py).
 * Nothing in the product path may load it.
 *
 * Everything is driven by the reference's own JSON schema and factory
 * (dmc/factory.h:15-98, dmc/main.cpp:188-245) so that the oracle sees exactly the objects an
 * input.json would create.
 *
 * Walker layout used by every entry point (same as the reference's Eigen column-major
 * state_t, dmc/traits.h:15): for each walker, species s = 0..S-1 contiguous, and inside a
 * species data[d*N_s + i]  (d = 0..2, i = particle).  One walker = 3*sum(N_s) doubles.
 */
#include "parameters.h"
#include "wavefunction/jastrows/jastrow.h"
#include "geometry.h"
#include "wavefunction/productWavefunction.h"
#include "tableDistances.h"
#include "potential.h"
#include "initializer.h"
#include "walkers.h"
#include "energy.h"
#include "estimators.h"
#include "dmcDriver.h"
#include "branching.h"
#include "moves/dmcMoves.h"
#include "wavefunction/jastrowWavefunctionOneBody.h"
#include "factory.h"
#include "wavefunction/jastrows/jastrowSquareWell.h"
#include "wavefunction/jastrows/jastrow_delta.h"
#include "wavefunction/jastrows/jastrowPoschTeller.h"
#include "wavefunction/jastrows/jastrowSpline.h"
#include "wavefunction/jastrows/jastrowGaussian.h"
#include "wavefunction/jastrows/jastrowGaussGauss.h"
#include "wavefunction/jastrows/jastrowHardSphereGauss.h"
#include "wavefunction/jastrows/jastrowDipolar.h"
#include "ptools.h"
#include "pairCorrelation.h"
#include "centerOfMassSquared.h"
#include "hyperRadius.h"
#include "trimerCorrelation.h"
#include "potentialFunctors.h"
#include "estimatorCollection.h"
#include "correlationEstimator.h"

#include <sstream>
#include <string>
#include <vector>
#include <memory>
#include <unistd.h>

namespace {

std::string g_err;

void registerAllOnce()
{
  static bool done = false;
  if (done) return;
  done = true;
  /* same registrations as dmc/main.cpp:188-216,236-243 minus the fermionic orbitals (out of scope); dipolar_rep runs on
     the Bessel stand-in of oracle/stubs/boost (oracle/bessel_k.h) */
  auto & f = getFactory();
  f.registerJastrow<gaussianJastrow>();
  f.registerJastrow<jastrowGaussGauss>();
  f.registerJastrow<jastrowSquareWell>();
  f.registerJastrow<jastrowSquareWellBoundState>();
  f.registerJastrow<jastrowPoschTeller>();
  f.registerJastrow<jastrowHardSphereGauss>();
  f.registerJastrow<jastrowHardSphere>();
  f.registerJastrow<jastrowSpline>();
  f.registerJastrow<jastrow_delta_phonons>();
  f.registerJastrow<jastrow_delta_in_trap>();
  f.registerJastrow<jastrow_delta_bound_state_phonons>();
  f.registerJastrow<jastrowDipolarRep>();
  f.registerObservable<pairCorrelation>();
  f.registerObservable<centerOfMassSquared>();
  f.registerObservable<trimerhyperRadius>();
  f.registerObservable<trimerCorrelation>();
  f.registerObservable<trimerCorrelationUnDis>();
  f.registerObservable<trimerCorrelationDis>();
  f.registerPotential<harmonicPotential>();
  f.registerPotential<squareWellPotential2b>();
  f.registerPotential<dipolarPotential2b>();
  f.registerPotential<poschTellerPotential2b>();
  f.registerPotentialFunctor<squareWell>();
  f.registerPotentialFunctor<potentialBarrier>();
}

struct model
{
  json_t j;
  std::vector<int> Ns;
  std::unique_ptr<geometryPBC> geo;
  std::vector<wavefunction*> waves;
  std::vector<potential*> pots;
  std::unique_ptr<productWavefunction> psi;
  std::unique_ptr<sumPotentials> pot;
  std::unique_ptr<energy> eO;
  int dofs = 0;

  explicit model(const char * jsonText)
  {
    registerAllOnce();
    std::istringstream ss(jsonText);
    ss >> j;
    auto lBox = j["lBox"].get<std::vector<real_t> >();
    Ns = j["N"].get<std::vector<int> >();
    geo.reset(new geometryPBC(lBox[0], lBox[1], lBox[2]));
    waves = getFactory().createWavefunctions(j["wavefunctions"], *geo);
    psi.reset(new productWavefunction(waves));
    if (j.find("potentials") != j.end()) pots = getFactory().createPotentials(j["potentials"], *geo);
    pot.reset(new sumPotentials(pots));
    eO.reset(new energy(pot.get()));
    for (int n : Ns) dofs += 3 * n;
  }
  ~model()
  {
    for (auto * w : waves) delete w;
    for (auto * p : pots) delete p;
  }

  states_t unpack(const double * R) const
  {
    states_t states;
    size_t off = 0;
    for (int n : Ns)
      {
        state_t s(n, 3);
        for (int d = 0; d < 3; d++)
          for (int i = 0; i < n; i++) s(i, d) = R[off + d * n + i];
        off += 3 * n;
        states.push_back(s);
      }
    return states;
  }
  void pack(const states_t & st, double * R) const
  {
    size_t off = 0;
    for (size_t s = 0; s < Ns.size(); s++)
      {
        int n = Ns[s];
        for (int d = 0; d < 3; d++)
          for (int i = 0; i < n; i++) R[off + d * n + i] = st[s](i, d);
        off += 3 * n;
      }
  }
};

template <class F> int guarded(F && f)
{
  try { f(); return 0; }
  catch (const std::exception & e) { g_err = e.what(); return 1; }
  catch (...) { g_err = "unknown exception"; return 2; }
}

} // namespace

/* Drives an estimatorCollection over a given genealogy the way dmcDriver does: clones are copies of the ancestor
 * walker (walkers.cpp:139-179), every walker of every event goes through estimatorCollection::accumulate
 * (dmcDriver.cpp:257-264: storers first, then estimators). */
static void run_genealogy(model & m, estimatorCollection & coll, const double * R, const int * n, const int * anc, int K)
{
  std::vector<dmcWalker> pop, next;
  size_t off = 0;
  for (int e = 0; e < K; e++)
    {
      next.clear();
      next.reserve(n[e]);
      for (int i = 0; i < n[e]; i++)
        {
          if (e == 0)
            {
              dmcWalker wk;
              wk.getStates() = m.unpack(R + (off + i) * m.dofs);
              initializer::registerDistances(wk.getTableDistances(), *m.psi);
              initializer::addDistances(wk, coll);
              coll.reserve(wk);
              next.push_back(wk);
            }
          else
            {
              next.push_back(pop.at(anc[off + i]));
              next.back().getStates() = m.unpack(R + (off + i) * m.dofs);
            }
          if (!updateForceGradientEnergy(next.back(), *m.psi, *m.eO)) throw std::runtime_error("invalid configuration");
        }
      for (auto & wk : next) coll.accumulate(wk, *m.psi);
      pop.swap(next);
      off += n[e];
    }
}

extern "C" {

const char * ref_last_error() { return g_err.c_str(); }

/* Per-configuration evaluation through updateForceGradientEnergy (dmc/walkers.cpp:46-61) after
 * initializer::initialize (dmc/initializer.cpp:75-82).  Additional distance tables needed by
 * potentials are registered from "measurements" exactly as dmcDriver::run does
 * (dmc/dmcDriver.cpp:175-187) when that key is present.
 * Outputs (any may be NULL): logpsi[nW], grad[nW*dofs], lap[nW], V[nW], E[nW], valid[nW]. */
int ref_eval(const char * jsonText, const double * R, int nW,
             double * logpsi, double * grad, double * lap, double * V, double * E, int * valid)
{
  return guarded([&] {
    model m(jsonText);
    for (int w = 0; w < nW; w++)
      {
        dmcWalker wk;
        auto st = m.unpack(R + (size_t)w * m.dofs);
        bool ok = initializer::initialize(wk, st, *m.psi, *m.eO);
        if (valid) valid[w] = ok ? 1 : 0;
        if (!ok)
          { /* constraint violated: the reference leaves derivatives untouched */
            if (logpsi) logpsi[w] = 0;
            if (lap) lap[w] = 0;
            if (V) V[w] = 0;
            if (E) E[w] = 0;
            if (grad) for (int k = 0; k < m.dofs; k++) grad[(size_t)w * m.dofs + k] = 0;
            continue;
          }
        if (logpsi) logpsi[w] = wk.getLogWave();
        if (lap) lap[w] = wk.getLaplacianLog();
        if (E) E[w] = wk.getEnergy();
        if (V) V[w] = (*m.pot)(wk);
        if (grad) m.pack(wk.getGradients(), grad + (size_t)w * m.dofs);
      }
  });
}

/* log Green-function ratio of driftDiffusionFirstOrder (dmc/moves/dmcMoves.cpp:32-57) for nW
 * (new, old) pairs; positions/gradients in the walker layout above. */
int ref_green_ratio(const char * jsonText, double timeStep, int nW,
                    const double * Rnew, const double * Fnew, const double * Rold, const double * Fold, double * q)
{
  return guarded([&] {
    model m(jsonText);
    driftDiffusionFirstOrder mover(timeStep);
    for (int w = 0; w < nW; w++)
      {
        dmcWalker a, b;
        a.getStates() = m.unpack(Rnew + (size_t)w * m.dofs);
        a.getGradients() = m.unpack(Fnew + (size_t)w * m.dofs);
        b.getStates() = m.unpack(Rold + (size_t)w * m.dofs);
        b.getGradients() = m.unpack(Fold + (size_t)w * m.dofs);
        q[w] = mover.transitionProbabilityRatio(a, b);
      }
  });
}

/* Drift part of driftDiffusionFirstOrder::move (dmc/moves/dmcMoves.cpp:12-30) followed by a
 * Gaussian displacement sqrt(timeStep)*xi with xi SUPPLIED by the caller in draw order
 * (species-major, particle-major, dimension-minor: dmc/moves/vmcMoves.cpp:22-33), so that the
 * move can be compared without sharing an RNG.  xi has dofs entries per walker in DRAW order. */
int ref_drift_diffuse(const char * jsonText, double timeStep, int nW,
                      const double * Rold, const double * Fold, const double * xi, double * Rnew)
{
  return guarded([&] {
    model m(jsonText);
    for (int w = 0; w < nW; w++)
      {
        auto R = m.unpack(Rold + (size_t)w * m.dofs);
        auto F = m.unpack(Fold + (size_t)w * m.dofs);
        const double * x = xi + (size_t)w * m.dofs;
        size_t k = 0;
        for (size_t s = 0; s < R.size(); s++)
          {
            R[s] = R[s] + F[s] * timeStep;            /* dmcMoves.cpp:23-26 */
            for (int i = 0; i < R[s].rows(); i++)
              for (int d = 0; d < 3; d++) R[s](i, d) = R[s](i, d) + x[k++] * sqrt(timeStep);   /* vmcMoves.cpp:31 */
          }
        m.pack(R, Rnew + (size_t)w * m.dofs);
      }
  });
}

/* Histogram estimators ("measurements" of kind pairCorrelation / trimerCorrelationUnDis / ...)
 * accumulated over nW configurations exactly as dmcDriver::accumulate does
 * (dmc/dmcDriver.cpp:257-264, dmc/estimatorCollection.cpp:32-43).
 * For measurement index iMeas (position inside j["measurements"] after the factory skipped
 * energy/forceEnergy kinds) returns sums[bins] and the weight.  The estimator constructors
 * append to <label>.dat in the CWD (dmc/estimators.cpp:17-22): we chdir to scratchDir first. */
int ref_histogram(const char * jsonText, const char * scratchDir, const double * R, int nW,
                  int iMeas, double * sums, int bins, double * weight)
{
  return guarded([&] {
    char cwd[4096];
    if (!getcwd(cwd, sizeof cwd)) throw std::runtime_error("getcwd");
    if (chdir(scratchDir) != 0) throw std::runtime_error("chdir scratch");
    try
      {
        model m(jsonText);
        auto ests = getFactory().createEstimators(m.j["measurements"]);
        estimatorCollection coll;
        for (auto * e : ests) coll.push_back(e);
        for (int w = 0; w < nW; w++)
          {
            dmcWalker wk;
            auto st = m.unpack(R + (size_t)w * m.dofs);
            wk.getStates() = st;
            initializer::registerDistances(wk.getTableDistances(), *m.psi);
            initializer::addDistances(wk, coll);
            if (!updateForceGradientEnergy(wk, *m.psi, *m.eO)) continue;
            coll.accumulate(wk, *m.psi);
          }
        if (auto * sc = dynamic_cast<realScalarEstimator*>(coll[iMeas].get()))
          { /* scalar observable: scalarAccumulator keeps sum and n private, average() = sum/n (accumulators.h:21-47) */
            if (bins != 1) throw std::runtime_error("scalar estimator: bins must be 1");
            auto & acc = sc->getAccumulator();
            *weight = (double)acc.getWeight();
            sums[0] = acc.getWeight() > 0 ? acc.average() * (double)acc.getWeight() : 0.;
            if (chdir(cwd) != 0) throw std::runtime_error("chdir back");
            return;
          }
        auto * h = dynamic_cast<realHistogramEstimator*>(coll[iMeas].get());
        if (!h) throw std::runtime_error("measurement is not a histogram estimator");
        auto & acc = h->getAccumulator();
        if ((int)acc.size() != bins) throw std::runtime_error("bins mismatch");
        for (int i = 0; i < bins; i++) sums[i] = acc.sums()(i);
        *weight = acc.weight();
      }
    catch (...) { if (chdir(cwd) != 0) {} throw; }
    if (chdir(cwd) != 0) throw std::runtime_error("chdir back");
  });
}

/* Forward walking through the reference's own classes over a GIVEN genealogy: K sampling events;
 * event e has n[e] walkers with configurations R (events concatenated) and anc[i] = index, in the
 * population of event e-1, of the walker this one descends from (ignored for e = 0).  Walkers are
 * copied the way branching copies them (whole dmcWalker, so the record rings follow:
 * walkers.cpp:139-179), then estimatorCollection::accumulate runs the storers
 * (realHistogramStorer::store, correlationEstimator.cpp:87-107) and the estimators
 * (realHistogramForwardWalkingEstimator::accumulate, :128-152) exactly as dmcDriver::accumulate does.
 * iEst indexes the vector estimatorFactory::create returns for j["measurements"]. */
int ref_forward_walking(const char * jsonText, const char * scratchDir, const double * R, const int * n,
                        const int * anc, int K, int iEst, double * sums, int bins, double * weight)
{
  return guarded([&] {
    char cwd[4096];
    if (!getcwd(cwd, sizeof cwd)) throw std::runtime_error("getcwd");
    if (chdir(scratchDir) != 0) throw std::runtime_error("chdir scratch");
    try
      {
        model m(jsonText);
        auto ests = getFactory().createEstimators(m.j["measurements"]);
        auto storers = getFactory().createStorers(m.j["measurements"]);
        estimatorCollection coll;
        for (auto * st : storers) coll.push_back(st);      /* main.cpp:411-420: storers, then estimators */
        for (auto * e : ests) coll.push_back(e);
        /* histogramAccumulator's default constructor leaves _weight uninitialised (accumulators.h:104) and
           realHistogramForwardWalkingEstimator only resizes it (correlationEstimator.cpp:156-158): the weight of the
           FIRST block of a reference run is garbage + count.  Every later block starts from
           estimatorCollection::clear() (dmcDriver.cpp:225) — that defined behaviour is what we pin. */
        coll.clear();
        run_genealogy(m, coll, R, n, anc, K);
        if (auto * sc = dynamic_cast<realScalarForwardWalkingEstimator*>(coll[iEst].get()))
          { /* scalar observable (correlationEstimator.cpp:36-51): accumulator += record; average() = sum/n */
            if (bins != 1) throw std::runtime_error("scalar forward-walking estimator: bins must be 1");
            auto & acc = sc->getAccumulator();
            *weight = (double)acc.getWeight();
            sums[0] = acc.getWeight() > 0 ? acc.average() * (double)acc.getWeight() : 0.;
            if (chdir(cwd) != 0) throw std::runtime_error("chdir back");
            return;
          }
        auto * h = dynamic_cast<realHistogramForwardWalkingEstimator*>(coll[iEst].get());
        if (!h) throw std::runtime_error("measurement is not a forward-walking histogram estimator");
        auto & acc = h->getAccumulator();
        if ((int)acc.size() != bins) throw std::runtime_error("bins mismatch");
        for (int i = 0; i < bins; i++) sums[i] = acc.sums()(i);
        *weight = acc.weight();
      }
    catch (...) { if (chdir(cwd) != 0) {} throw; }
    if (chdir(cwd) != 0) throw std::runtime_error("chdir back");
  });
}

/* forceEnergy::operator() (dmc/energy.cpp:37-61) of the reference on nW configurations: out[nW]. */
int ref_force_energy(const char * jsonText, const double * R, int nW, double * out)
{
  return guarded([&] {
    model m(jsonText);
    forceEnergy fe(m.pot.get());
    for (int w = 0; w < nW; w++)
      {
        dmcWalker wk;
        auto st = m.unpack(R + (size_t)w * m.dofs);
        if (!initializer::initialize(wk, st, *m.psi, *m.eO)) throw std::runtime_error("invalid configuration");
        out[w] = fe(wk, *m.psi);
      }
  });
}

/* "superfluidFraction" measurements: the reference's centerOfMassStorer + superfluidFractionEstimator
 * (correlationEstimator.cpp:187-357) created by its own factories (estimatorsFactory.cpp:80-116,
 * estimatorsFactory.h:90-131) and driven over a genealogy; sums/weights of estimator iEst. */
int ref_superfluid_fraction(const char * jsonText, const char * scratchDir, const double * R, const int * n,
                            const int * anc, int K, int iEst, double * sums, double * weights, int bins)
{
  return guarded([&] {
    char cwd[4096];
    if (!getcwd(cwd, sizeof cwd)) throw std::runtime_error("getcwd");
    if (chdir(scratchDir) != 0) throw std::runtime_error("chdir scratch");
    try
      {
        model m(jsonText);
        auto ests = getFactory().createEstimators(m.j["measurements"]);
        auto storers = getFactory().createStorers(m.j["measurements"]);
        estimatorCollection coll;
        for (auto * st : storers) coll.push_back(st);
        for (auto * e : ests) coll.push_back(e);
        coll.clear();
        run_genealogy(m, coll, R, n, anc, K);
        auto * sf = dynamic_cast<superfluidFractionEstimator*>(coll[iEst].get());
        if (!sf) throw std::runtime_error("measurement is not a superfluidFraction estimator");
        auto & acc = sf->getAccumulator();
        if ((int)acc.size() != bins) throw std::runtime_error("bins mismatch");
        for (int i = 0; i < bins; i++) { sums[i] = acc.sums()(i); weights[i] = acc.weights()(i); }
      }
    catch (...) { if (chdir(cwd) != 0) {} throw; }
    if (chdir(cwd) != 0) throw std::runtime_error("chdir back");
  });
}

/* branchingControl (dmc/branching.cpp): E_T from setEnergyShift(walkers with energies Eold),
 * then nDescendants for each (Enew[i], Eold[i]) with a fresh std::ranlux48(seed), then the
 * in-place branch() replayed with the same seed (as tst/dmc-test.cpp:15-83 does).
 * Outputs: energyShift, nDesc[n], u[n] (the uniform each walker consumed, recovered by
 * replaying the generator), newPop and the energies of the branched population (capacity capOut). */
int ref_branch(double timeStep, double meanWalkers, double deltaWalkers, int seed, int n,
               const double * Enew, const double * Eold,
               double * energyShift, int * nDesc, double * u, int * newPop, double * Eafter, int capOut,
               double * energyShiftAfter)
{
  return guarded([&] {
    walkerContainer<dmcWalker> oldW, newW;
    dmcWalker proto;
    oldW.resize(n, proto);
    newW.resize(n, proto);
    for (int i = 0; i < n; i++) { oldW[i].getEnergy() = Eold[i]; newW[i].getEnergy() = Enew[i]; }
    branchingControl bc(timeStep, meanWalkers, deltaWalkers);
    bc.setEnergyShift(oldW);
    /* recover E_T: nDescendants with u -> we cannot read energyShift directly (private), so
       recompute it the way setEnergyShift does and cross-check through nDescendants below */
    double eT = 0;
    for (int i = 0; i < n; i++) eT += Eold[i];
    eT /= n;
    if (n > (size_t(meanWalkers) + size_t(deltaWalkers))) eT += 1. / timeStep * log(size_t(meanWalkers) * 1. / (size_t(meanWalkers) + size_t(deltaWalkers)));
    else if (n < (size_t(meanWalkers) - size_t(deltaWalkers))) eT += 1. / timeStep * log(size_t(meanWalkers) * 1. / (size_t(meanWalkers) - size_t(deltaWalkers)));
    *energyShift = eT;
    {
      randomGenerator_t r1(seed), r2(seed);
      std::uniform_real_distribution<real_t> dis(0., 1.);
      for (int i = 0; i < n; i++)
        {
          nDesc[i] = bc.nDescendants(newW[i], oldW[i], r1);
          u[i] = dis(r2);
        }
    }
    randomGenerator_t r3(seed);
    bc.branch(newW, oldW, r3);
    *newPop = (int)newW.size();
    for (int i = 0; i < (int)newW.size() && i < capOut; i++) Eafter[i] = newW[i].getEnergy();
    /* E_T of the branched population (dmcDriver.cpp:141-143) recovered through one more
       nDescendants call is not needed: report the plain formula inputs instead */
    double s = 0;
    for (size_t i = 0; i < newW.size(); i++) s += newW[i].getEnergy();
    double pop = (double)newW.size();
    double e2 = s / pop;
    if (newW.size() > (size_t(meanWalkers) + size_t(deltaWalkers))) e2 += 1. / timeStep * log(size_t(meanWalkers) * 1. / (size_t(meanWalkers) + size_t(deltaWalkers)));
    else if (newW.size() < (size_t(meanWalkers) - size_t(deltaWalkers))) e2 += 1. / timeStep * log(size_t(meanWalkers) * 1. / (size_t(meanWalkers) - size_t(deltaWalkers)));
    *energyShiftAfter = e2;
  });
}

/* Alias-method load-balance planner (dmc/ptools.cpp:93-161).  populations[nRanks] in; for
 * every rank r: sources[r] = rank it receives from, amounts[r] = how many walkers. */
int ref_plan(const int * populations, int nRanks, int * sources, int * amounts)
{
  return guarded([&] {
    std::vector<int> pops(populations, populations + nRanks), perm, src, amt;
    std::vector<std::vector<int> > dest;
    pTools::determineLoadBalanceComunicationsAliasMethod(pops, perm, src, dest, amt);
    for (int r = 0; r < nRanks; r++) { sources[r] = src[r]; amounts[r] = amt[r]; }
  });
}

/* Radial Jastrow functor tabulation through the wavefunction factory: evaluates u,u',u'' of
 * the "jastrow" object of wavefunctions[iWave] is not reachable generically (J is private), so
 * the tabulation is obtained from a 2-particle system instead: see tests. */

} // extern "C"
