// qmc4host.h — C++ host side of the drop-in: the reference's factory / wavefunction / potential /
// estimator / dmcDriver API for the DMC path, re-implemented over the C ABI of include/qmc4_b200.h.
//
// Same class names, registration surface, id strings, JSON keys, defaults, block report and output
// files as the reference, so an existing input.json (and the user's habit of registering types on
// getFactory()) keeps working:
//   factory / getFactory()            dmc/factory.h:15-98, dmc/factory.cpp:3-13
//   createId                          dmc/abstractFactory.cpp:3-20
//   wavefunction, productWavefunction dmc/wavefunction/wavefunction.h:34-81, productWavefunction.h
//   potential, sumPotentials          dmc/potential.h:12-48
//   estimators                        dmc/estimators.{h,cpp}, dmc/estimatorCollection.h
//   dmcDriver                         dmc/dmcDriver.{h,cpp}, dmc/driver.{h,cpp}
// What changes: a registered type no longer carries CPU virtuals that evaluate physics; it carries a
// *device term descriptor* (emit()) that issues the matching qmc_add_* call.  All numerics run in the
// CUDA library.
#pragma once
#include "../../include/qmc4_b200.h"
#include "input.h"
#include "qmcExceptions.h"

#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <vector>

using real_t = double;
using state_t = std::vector<real_t>;      // one species, column-major N x 3 (data[d*N+i], dmc/traits.h:15)
using states_t = std::vector<state_t>;
inline int getN(const state_t &s) { return (int)s.size() / 3; }
constexpr int getDimensions() { return 3; }

class geometryPBC {
 public:
  geometryPBC(real_t lx, real_t ly, real_t lz) : lBox{lx, ly, lz} {}
  explicit geometryPBC(real_t l) : geometryPBC(l, l, l) {}
  real_t getLBox(int i) const { return lBox[i]; }

 private:
  real_t lBox[3];
};
using geometry_t = geometryPBC;

std::string createId(const json_t &j);

// ---- radial Jastrow functors: descriptors with the reference's names and JSON keys -------------------
struct jastrowBase {
  int kind = 0;
  std::vector<real_t> params;
};
#define QMC_JASTROW(cls, idname, kindenum, ...)                         \
  struct cls : jastrowBase {                                            \
    static std::string name() { return idname; }                        \
    explicit cls(const json_t &j)                                       \
    {                                                                   \
      kind = kindenum;                                                  \
      for (const char *k : {__VA_ARGS__}) params.push_back(j[k].get<real_t>()); \
    }                                                                   \
  }
QMC_JASTROW(gaussianJastrow, "gaussian", QMC_J_GAUSSIAN, "alpha");                                   // jastrowGaussian.h
QMC_JASTROW(jastrowGaussGauss, "gaussiangaussian", QMC_J_GAUSSGAUSS, "r0", "u0");                    // jastrowGaussGauss.cpp:12-17
QMC_JASTROW(jastrowSquareWell, "squareWell", QMC_J_SQUAREWELL, "V0", "R0", "Rm", "alpha", "aInverse", "cut_off");   // jastrowSquareWell.cpp:42-44
QMC_JASTROW(jastrowSquareWellBoundState, "squareWellBoundState", QMC_J_SQUAREWELL_BOUND, "V0", "R0", "Rm", "cut_off", "aInverse", "eb", "alpha");
QMC_JASTROW(jastrowPoschTeller, "poschTeller", QMC_J_POSCHTELLER, "R0", "C", "alpha", "Rm", "cut_off");
QMC_JASTROW(jastrowHardSphereGauss, "hardSphereGauss", QMC_J_HARDSPHEREGAUSS, "a", "Rm", "D");
QMC_JASTROW(jastrowHardSphere, "hardSphere", QMC_J_HARDSPHERE, "a", "k", "cut_off");
QMC_JASTROW(jastrow_delta_in_trap, "delta_in_trap", QMC_J_DELTA_IN_TRAP, "a");                        // jastrow_delta.cpp:71-74
QMC_JASTROW(jastrowDipolarRep, "dipolar_rep", QMC_J_DIPOLAR_REP, "D", "cut_off", "matching_point", "alpha", "C");   // jastrowDipolar.cpp:3-13
#undef QMC_JASTROW
struct jastrow_delta_phonons : jastrowBase {                                                          // jastrow_delta.cpp:5-39
  static std::string name() { return "delta_phonons"; }
  explicit jastrow_delta_phonons(const json_t &j)
  {
    kind = QMC_J_DELTA_PHONONS;
    for (const char *k : {"k", "delta", "beta", "z", "cut_off"}) params.push_back(j[k].get<real_t>());
    try
      {
        (void)j["g"].get<real_t>();        /* parsed for process(), which the reference never calls */
      }
    catch (invalidType &)
      {
        if (j["g"].get<std::string>() != "inf") throw invalidInput("g is not a number and is not infinite");
      }
  }
};
struct jastrow_delta_bound_state_phonons : jastrowBase {                                              // jastrow_delta.cpp:76-86
  static std::string name() { return "delta_bound_state_phonons"; }
  explicit jastrow_delta_bound_state_phonons(const json_t &j)
  {
    kind = QMC_J_DELTA_BOUND_STATE_PHONONS;
    const real_t xI = j["xI"].get<real_t>(), g = j["g"].get<real_t>(), cut = j["cut_off"].get<real_t>(), k = j["k"].get<real_t>(),
                 A = j["A"].get<real_t>(), beta = j["beta"].get<real_t>();
    (void)g;
    params = {xI, k, A, beta, cut};
  }
};
struct jastrowSpline : jastrowBase {                                                                  // jastrowSpline.cpp:20-22
  static std::string name() { return "bSpline"; }
  explicit jastrowSpline(const json_t &j)
  {
    kind = QMC_J_BSPLINE;
    params = {j["stepSize"].get<real_t>(), j["valueRight"].get<real_t>(), j["derivativeRight"].get<real_t>()};
    for (real_t c : j["coefficients"].get<std::vector<real_t> >()) params.push_back(c);
  }
};

// ---- wavefunction terms -----------------------------------------------------------------------------
class wavefunction {
 public:
  explicit wavefunction(const geometry_t &geo_) : geo(&geo_) {}
  virtual ~wavefunction() {}
  const geometry_t &getGeometry() const { return *geo; }
  virtual std::vector<int> sets() const = 0;
  virtual void emit(qmc_ctx *ctx) const = 0;       // device term descriptor
  // rows of print() = J.print(0, lBox/2, n) (jastrowWavefunctionTwoBody.h:81-83,169: 10000; ...ThreeBody.h:140-144,286-290: 100000;
  // one-body terms keep wavefunction::print() = "" (wavefunction.h:64))
  virtual size_t printPoints() const { return 0; }
 protected:
  void add(qmc_ctx *ctx, int wfKind, const std::vector<int> &s, const jastrowBase &J, real_t hs) const;

 private:
  const geometry_t *geo;
};

template <class jastrow_t> class jastrowOneBodyWavefunction : public wavefunction {
 public:
  jastrowOneBodyWavefunction(const json_t &j, const geometry_t &geo) : wavefunction(geo), J(j["jastrow"]), setA(j["set"].get<int>()) {}
  static std::string name() { return "jastrow1b/" + jastrow_t::name(); }
  std::vector<int> sets() const override { return {setA}; }
  void emit(qmc_ctx *ctx) const override { add(ctx, QMC_WF_JASTROW1B, {setA}, J, -1); }

 private:
  jastrow_t J;
  int setA;
};

template <class jastrow_t> class jastrowTwoBodyWavefunctionUndistinguishable : public wavefunction {
 public:
  jastrowTwoBodyWavefunctionUndistinguishable(const json_t &j, const geometry_t &geo)
      : wavefunction(geo), J(j["jastrow"]), setA(j["sets"][0].get<int>()), setB(j["sets"][1].get<int>())
  {
    if (setA != setB) throw invalidInput("setA != setB in undistinguishable two body jastrow");
  }
  static std::string name() { return "jastrow2bUnDis/" + jastrow_t::name(); }
  std::vector<int> sets() const override { return {setA, setB}; }
  void emit(qmc_ctx *ctx) const override { add(ctx, QMC_WF_JASTROW2B_UNDIS, {setA, setB}, J, -1); }
  size_t printPoints() const override { return 10000; }

 private:
  jastrow_t J;
  int setA, setB;
};

template <class jastrow_t> class jastrowTwoBodyWavefunctionDistinguishable : public wavefunction {
 public:
  jastrowTwoBodyWavefunctionDistinguishable(const json_t &j, const geometry_t &geo)
      : wavefunction(geo), J(j["jastrow"]), setA(j["sets"][0].get<int>()), setB(j["sets"][1].get<int>())
  {
    if (setA == setB) throw invalidInput("setA == setB in distinguishable two body jastrow");
  }
  static std::string name() { return "jastrow2bDis/" + jastrow_t::name(); }
  std::vector<int> sets() const override { return {setA, setB}; }
  void emit(qmc_ctx *ctx) const override { add(ctx, QMC_WF_JASTROW2B_DIS, {setA, setB}, J, -1); }
  size_t printPoints() const override { return 10000; }

 private:
  jastrow_t J;
  int setA, setB;
};

template <class jastrow_t> class jastrowThreeBodyWavefunctionUnDistinguishable : public wavefunction {
 public:
  jastrowThreeBodyWavefunctionUnDistinguishable(const json_t &j, const geometry_t &geo)
      : wavefunction(geo), J(j["jastrow"]), setA(j["sets"][0].get<int>()), hs(j.contains("hardSphereRadius") ? j["hardSphereRadius"].get<real_t>() : -1)
  {
  }
  static std::string name() { return "jastrow3bUnDis/" + jastrow_t::name(); }
  std::vector<int> sets() const override { return {setA, setA}; }
  void emit(qmc_ctx *ctx) const override { add(ctx, QMC_WF_JASTROW3B_UNDIS, {setA}, J, hs); }
  size_t printPoints() const override { return 100000; }

 private:
  jastrow_t J;
  int setA;
  real_t hs;
};

template <class jastrow_t> class jastrowThreeBodyWavefunctionDistinguishable : public wavefunction {   // jastrowWavefunctionThreeBody.h:164-305
 public:
  jastrowThreeBodyWavefunctionDistinguishable(const json_t &j, const geometry_t &geo)
      : wavefunction(geo), J(j["jastrow"]), setA(j["sets"][0].get<int>()), setB(j["sets"][1].get<int>()), setC(j["sets"][2].get<int>()),
        hs(j.contains("hardSphereRadius") ? j["hardSphereRadius"].get<real_t>() : -1)
  {
    if (setA == setB || setB == setC || setA == setC) throw invalidInput("Sets in three body wavefunction should be distinguishible");
  }
  static std::string name() { return "jastrow3bDis/" + jastrow_t::name(); }
  std::vector<int> sets() const override { return {setA, setB, setC}; }
  void emit(qmc_ctx *ctx) const override { add(ctx, QMC_WF_JASTROW3B_DIS, {setA, setB, setC}, J, hs); }
  size_t printPoints() const override { return 100000; }

 private:
  jastrow_t J;
  int setA, setB, setC;
  real_t hs;
};

class productWavefunction {
 public:
  productWavefunction() {}
  explicit productWavefunction(std::vector<wavefunction *> waves) : _logWaves(waves) {}
  void add(wavefunction *w) { _logWaves.push_back(w); }
  const std::vector<wavefunction *> &waves() const { return _logWaves; }
  size_t size() const { return _logWaves.size(); }
  const geometry_t &getGeometry() const { return _logWaves.at(0)->getGeometry(); }

 private:
  std::vector<wavefunction *> _logWaves;      // not owned (dmc/wavefunction/productWavefunction.h:16-19)
};
using wavefunction_t = productWavefunction;

// ---- potentials ---------------------------------------------------------------------------------------
class potential {
 public:
  explicit potential(const geometry_t &geo_) : geo(&geo_) {}
  virtual ~potential() {}
  virtual std::vector<int> sets() const = 0;
  virtual void emit(qmc_ctx *ctx) const = 0;
 protected:
  void add(qmc_ctx *ctx, int kind, const std::vector<int> &s, const std::vector<real_t> &p) const;

 private:
  const geometry_t *geo;
};

class harmonicPotential : public potential {
 public:
  harmonicPotential(const json_t &j, const geometry_t &geo) : potential(geo), omega(j["omega"].get<real_t>()), setA(j["set"].get<int>()) {}
  static std::string name() { return "harmonic1b"; }
  std::vector<int> sets() const override { return {setA}; }
  void emit(qmc_ctx *ctx) const override { add(ctx, QMC_POT_HARMONIC1B, {setA}, {omega}); }

 private:
  real_t omega;
  int setA;
};
#define QMC_POT2B(cls, idname, kindenum, ...)                                                                  \
  class cls : public potential {                                                                               \
   public:                                                                                                     \
    cls(const json_t &j, const geometry_t &geo) : potential(geo), s{j["sets"][0].get<int>(), j["sets"][1].get<int>()} \
    {                                                                                                          \
      for (const char *k : {__VA_ARGS__}) p.push_back(j[k].get<real_t>());                                     \
    }                                                                                                          \
    static std::string name() { return idname; }                                                               \
    std::vector<int> sets() const override { return s; }                                                       \
    void emit(qmc_ctx *ctx) const override { add(ctx, kindenum, s, p); }                                       \
   private:                                                                                                    \
    std::vector<int> s;                                                                                        \
    std::vector<real_t> p;                                                                                     \
  }
QMC_POT2B(squareWellPotential2b, "squareWell2b", QMC_POT_SQUAREWELL2B, "V0", "R0");      // potential.cpp:54-55
QMC_POT2B(dipolarPotential2b, "dipolar2b", QMC_POT_DIPOLAR2B, "D");                      // potential.cpp:70
QMC_POT2B(poschTellerPotential2b, "poschTeller2b", QMC_POT_POSCHTELLER2B, "R0");         // potential.cpp:127-131
#undef QMC_POT2B

struct squareWell {        // potentialFunctors.h:5-19
  explicit squareWell(const json_t &j) : R0(j["R0"].get<real_t>()), V0(j["V0"].get<real_t>()) {}
  static std::string name() { return "squareWell"; }
  static int kind2b() { return QMC_POT_TWOBODY_SQUAREWELL; }
  static int kind3b() { return QMC_POT_THREEBODY_UNDIS_SQUAREWELL; }
  static int kind3bDis() { return QMC_POT_THREEBODY_DIS_SQUAREWELL; }
  real_t R0, V0;
};
struct potentialBarrier {  // potentialFunctors.h:22-36
  explicit potentialBarrier(const json_t &j) : R0(j["R0"].get<real_t>()), V0(j["V0"].get<real_t>()) {}
  static std::string name() { return "potentialBarrier"; }
  static int kind2b() { return QMC_POT_TWOBODY_BARRIER; }
  static int kind3b() { return QMC_POT_THREEBODY_UNDIS_BARRIER; }
  static int kind3bDis() { return QMC_POT_THREEBODY_DIS_BARRIER; }
  real_t R0, V0;
};
template <class functor_t> class potentialTwoBody : public potential {      // potentialTwoBody.h
 public:
  potentialTwoBody(const json_t &j, const geometry_t &geo) : potential(geo), V(j["functor"]), setA(j["sets"][0].get<int>()), setB(j["sets"][1].get<int>()) {}
  static std::string name() { return "potentialTwoBody/" + functor_t::name(); }
  std::vector<int> sets() const override { return {setA, setB}; }
  void emit(qmc_ctx *ctx) const override { add(ctx, functor_t::kind2b(), {setA, setB}, {V.R0, V.V0}); }

 private:
  functor_t V;
  int setA, setB;
};
template <class functor_t> class potentialThreeBodyUnDis : public potential {   // potentialThreeBody.h:8-61
 public:
  potentialThreeBodyUnDis(const json_t &j, const geometry_t &geo) : potential(geo), V(j["functor"]), setA(j["sets"][0].get<int>()) {}
  static std::string name() { return "potentialThreeBodyUnDis/" + functor_t::name(); }
  std::vector<int> sets() const override { return {setA, setA, setA}; }
  void emit(qmc_ctx *ctx) const override { add(ctx, functor_t::kind3b(), {setA, setA, setA}, {V.R0, V.V0}); }

 private:
  functor_t V;
  int setA;
};

template <class functor_t> class potentialThreeBodyDis : public potential {     // potentialThreeBody.h:60-125
 public:
  potentialThreeBodyDis(const json_t &j, const geometry_t &geo)
      : potential(geo), V(j["functor"]), setA(j["sets"][0].get<int>()), setB(j["sets"][1].get<int>()), setC(j["sets"][2].get<int>())
  {
    if (setA == setB || setB == setC || setA == setC) throw invalidInput("Sets in three body wavefunction should be distinguishible");   /* sic, :70 */
  }
  static std::string name() { return "potentialThreeBodyDis/" + functor_t::name(); }
  std::vector<int> sets() const override { return {setA, setB, setC}; }
  void emit(qmc_ctx *ctx) const override { add(ctx, functor_t::kind3bDis(), {setA, setB, setC}, {V.R0, V.V0}); }

 private:
  functor_t V;
  int setA, setB, setC;
};

class sumPotentials {
 public:
  explicit sumPotentials(std::vector<potential *> p) : _potentials(p) {}
  const std::vector<potential *> &potentials() const { return _potentials; }
  size_t size() const { return _potentials.size(); }

 private:
  std::vector<potential *> _potentials;
};
using potential_t = sumPotentials;

// ---- observables / estimators -----------------------------------------------------------------------
struct observableBase {
  virtual ~observableBase() {}
  virtual int emit(qmc_ctx *ctx, int bins, real_t minx, real_t maxx) const = 0;   // returns the device estimator index
  virtual std::vector<int> sets() const = 0;
  virtual bool isVector() const { return false; }
  virtual bool isScalar() const { return false; }        // realScalarObservable (observables.h:10-30)
  virtual std::vector<real_t> x(int) const { return {}; }
};
// observables over three sets: trimerCorrelation (set C = one impurity, trimerCorrelation.cpp:9-78), trimerCorrelationDis
// (:163-227) are histograms; trimerhyperRadius (hyperRadius.cpp) is a scalar
template <int EST_KIND, bool SCALAR> class threeSetObservable : public observableBase {
 public:
  explicit threeSetObservable(const json_t &j) : setA(j["sets"][0].get<int>()), setB(j["sets"][1].get<int>()), setC(j["sets"][2].get<int>()) {}
  std::vector<int> sets() const override { return {setA, setB, setC}; }
  bool isScalar() const override { return SCALAR; }
  int emit(qmc_ctx *ctx, int bins, real_t minx, real_t maxx) const override;

 private:
  int setA, setB, setC;
};
struct trimerCorrelation : threeSetObservable<QMC_EST_TRIMERCORRELATION, false> {
  using threeSetObservable::threeSetObservable;
  static std::string name() { return "trimerCorrelation"; }
};
struct trimerCorrelationDis : threeSetObservable<QMC_EST_TRIMERCORRELATION_DIS, false> {
  using threeSetObservable::threeSetObservable;
  static std::string name() { return "trimerCorrelationDis"; }
};
struct trimerhyperRadius : threeSetObservable<QMC_EST_TRIMERHYPERRADIUS, true> {
  using threeSetObservable::threeSetObservable;
  static std::string name() { return "trimerhyperRadius"; }
};
class centerOfMassSquared : public observableBase {             // centerOfMassSquared.cpp:6 ("set", not "sets")
 public:
  explicit centerOfMassSquared(const json_t &j) : setA(j["set"].get<int>()) {}
  static std::string name() { return "centerOfMassSquared"; }
  std::vector<int> sets() const override { return {setA}; }
  bool isScalar() const override { return true; }
  int emit(qmc_ctx *ctx, int bins, real_t minx, real_t maxx) const override;

 private:
  int setA;
};
class pairCorrelation : public observableBase {                 // pairCorrelation.cpp:9-15
 public:
  explicit pairCorrelation(const json_t &j) : setA(j["sets"][0].get<int>()), setB(j["sets"][1].get<int>()) {}
  static std::string name() { return "pairCorrelation"; }
  std::vector<int> sets() const override { return {setA, setB}; }
  int emit(qmc_ctx *ctx, int bins, real_t minx, real_t maxx) const override;

 private:
  int setA, setB;
};
class trimerCorrelationUnDis : public observableBase {          // trimerCorrelation.cpp:118-126
 public:
  explicit trimerCorrelationUnDis(const json_t &j) : setA(j["sets"][0].get<int>()) {}
  static std::string name() { return "trimerCorrelationUnDis"; }
  std::vector<int> sets() const override { return {setA, setA}; }
  int emit(qmc_ctx *ctx, int bins, real_t minx, real_t maxx) const override;

 private:
  int setA;
};
class structureFactor : public observableBase {                 // structureFactor.cpp:42-54 (q along x)
 public:
  explicit structureFactor(const json_t &j)
      : setA(j["sets"][0].get<int>()), setB(j["sets"][1].get<int>()), bins(j["bins"].get<int>()), lBox(j["lBox"].get<real_t>()),
        maxq(j["maxq"].get<real_t>()), channel(j["channel"].get<int>())
  {
  }
  static std::string name() { return "structureFactor"; }
  std::vector<int> sets() const override { return {setA, setB}; }
  bool isVector() const override { return true; }
  std::vector<real_t> x(int n) const override;
  int emit(qmc_ctx *ctx, int bins, real_t minx, real_t maxx) const override;

 private:
  int setA, setB, bins;
  real_t lBox, maxq;
  int channel;
};

class estimatorBase {                                           // estimators.cpp:17-33
 public:
  explicit estimatorBase(const std::string &label);
  virtual ~estimatorBase() {}
  const std::string &label() const { return _label; }
  virtual void registerOn(qmc_ctx *) {}
  virtual void collect(qmc_ctx *, const qmc_block_result &) = 0;     // fetch block data from the device
  virtual void write(std::ostream &stream) = 0;
  virtual void dump();
  virtual std::vector<int> sets() const { return {}; }

 protected:
  std::string _label, filename, pendingHeader;
  std::fstream f;
  void headerIfEmpty(const std::string &h);
};
class realScalarEstimator : public estimatorBase {              // the implicit "energy" estimator (dmcDriver.cpp:232-245)
 public:
  explicit realScalarEstimator(const std::string &label);
  void collect(qmc_ctx *, const qmc_block_result &r) override { sum = r.energy_sum; n = r.energy_count; }
  void write(std::ostream &stream) override;
  real_t average() const { return sum / n; }

 private:
  real_t sum = 0, n = 0;
};
class realScalarObservableEstimator : public estimatorBase {    // realScalarEstimator(ob, j) (estimators.cpp:48): reads "label" only
 public:
  realScalarObservableEstimator(observableBase *ob, const json_t &j);
  void registerOn(qmc_ctx *ctx) override { index = ob->emit(ctx, 1, 0, 1); }
  void collect(qmc_ctx *ctx, const qmc_block_result &) override;
  void write(std::ostream &stream) override;
  std::vector<int> sets() const override { return ob->sets(); }

 private:
  std::unique_ptr<observableBase> ob;
  real_t sum = 0, n = 0;
  int index = -1;
};
class realHistogramEstimator : public estimatorBase {           // estimators.cpp:50-87
 public:
  realHistogramEstimator(observableBase *ob, const json_t &j);
  void registerOn(qmc_ctx *ctx) override { index = ob->emit(ctx, (int)sums.size(), minx, maxx); }
  void collect(qmc_ctx *ctx, const qmc_block_result &) override;
  void write(std::ostream &stream) override;
  std::vector<int> sets() const override { return ob->sets(); }

 private:
  std::unique_ptr<observableBase> ob;
  std::vector<real_t> sums, x;
  real_t weight = 0, minx, maxx;
  int index = -1;
  bool vectorStyle;
};

// Forward walking (correlationEstimator.h:129-145, estimatorsFactory.cpp:47-74): a measurement with
// "forwardWalkingSteps": [..] yields one estimator "<label>_fw<steps>" per entry instead of the plain one,
// fed by the storer that "recordSteps" creates (estimatorsFactory.h:97-131).  On the device the storer is
// the record ring of qmc_add_forward_walking; the siblings of one measurement share its registration.
struct forwardWalkingGroup {
  std::unique_ptr<observableBase> ob;
  int bins = 0, recordSteps = 0, first = -1;
  real_t minx = 0, maxx = 0;
  std::vector<int> steps;
  void ensureRegistered(qmc_ctx *ctx);
};
class realHistogramForwardWalkingEstimator : public estimatorBase {
 public:
  realHistogramForwardWalkingEstimator(std::shared_ptr<forwardWalkingGroup> group, int k, const std::string &label);
  void registerOn(qmc_ctx *ctx) override { group->ensureRegistered(ctx); }
  void collect(qmc_ctx *ctx, const qmc_block_result &) override;
  void write(std::ostream &stream) override;
  std::vector<int> sets() const override { return {}; }

 private:
  std::shared_ptr<forwardWalkingGroup> group;
  int k;
  std::vector<real_t> sums;
  real_t weight = 0;
};

class realScalarForwardWalkingEstimator : public estimatorBase {   // correlationEstimator.cpp:36-80
 public:
  realScalarForwardWalkingEstimator(std::shared_ptr<forwardWalkingGroup> group, int k, const std::string &label);
  void registerOn(qmc_ctx *ctx) override { group->ensureRegistered(ctx); }
  void collect(qmc_ctx *ctx, const qmc_block_result &) override;
  void write(std::ostream &stream) override;

 private:
  std::shared_ptr<forwardWalkingGroup> group;
  int k;
  real_t sum = 0, n = 0;
};

// main.cpp:270-273: realScalarEstimator("forceEnergy", new forceEnergy(&pot)) — always constructed (so forceEnergy.dat always
// gets its header on rank 0), attached to the driver only for vmc (:318-335) and svmc (:395-408) when "measurements" has a
// {"kind": "forceEnergy"} entry.  forceEnergy::operator() (energy.cpp:37-61) runs on the device (qmc_enable_force_energy).
class forceEnergyEstimator : public estimatorBase {
 public:
  forceEnergyEstimator();
  void writeHeaderNow();                                        // estimators.cpp:35-46 (rank 0, at construction)
  void registerOn(qmc_ctx *ctx) override;
  void collect(qmc_ctx *ctx, const qmc_block_result &) override;
  void write(std::ostream &stream) override;

 private:
  real_t sum = 0, n = 0;
};

// superfluidFractionEstimator (correlationEstimator.h:162-180, .cpp:272-357): winding estimator over the centre-of-mass
// records the walkers carry; on the device the records live per slot on the forward-walking event ring
// (qmc_add_superfluid_fraction).  Sets are species indices of the input (the factory's remapping through the storer's
// set list, estimatorsFactory.cpp:86-108, happens inside the library).
class superfluidFractionEstimator : public estimatorBase {
 public:
  explicit superfluidFractionEstimator(const json_t &j);
  void registerOn(qmc_ctx *ctx) override;
  void collect(qmc_ctx *ctx, const qmc_block_result &) override;
  void write(std::ostream &stream) override;
  std::vector<int> sets() const override { return {setA, setB}; }

 private:
  int setA, setB, index = -1;
  std::vector<real_t> sums, weights;
};

// ---- factory (dmc/factory.h) -------------------------------------------------------------------------
class factory {
 public:
  using waveCreator = std::function<wavefunction *(const json_t &, const geometry_t &)>;
  using potCreator = std::function<potential *(const json_t &, const geometry_t &)>;
  using obCreator = std::function<observableBase *(const json_t &)>;

  template <class wave_t> void registerWavefunction()
  {
    waves[wave_t::name()] = [](const json_t &j, const geometry_t &g) -> wavefunction * { return new wave_t(j, g); };
  }
  template <class pot_t> void registerPotential()
  {
    pots[pot_t::name()] = [](const json_t &j, const geometry_t &g) -> potential * { return new pot_t(j, g); };
  }
  template <class functor_t> void registerPotentialFunctor()
  {
    registerPotential<potentialTwoBody<functor_t> >();
    registerPotential<potentialThreeBodyUnDis<functor_t> >();
    registerPotential<potentialThreeBodyDis<functor_t> >();
  }
  template <class jastrow_t> void registerJastrow()
  { /* one call registers the Jastrow for every wavefunction template (factory.h:50-60) */
    registerWavefunction<jastrowOneBodyWavefunction<jastrow_t> >();
    registerWavefunction<jastrowTwoBodyWavefunctionUndistinguishable<jastrow_t> >();
    registerWavefunction<jastrowTwoBodyWavefunctionDistinguishable<jastrow_t> >();
    registerWavefunction<jastrowThreeBodyWavefunctionUnDistinguishable<jastrow_t> >();
    registerWavefunction<jastrowThreeBodyWavefunctionDistinguishable<jastrow_t> >();
  }
  template <class ob_t> void registerObservable()
  {
    obs[ob_t::name()] = [](const json_t &j) -> observableBase * { return new ob_t(j); };
  }
  std::vector<wavefunction *> createWavefunctions(const json_t &j, const geometry_t &geo);
  std::vector<potential *> createPotentials(const json_t &j, const geometry_t &geo);
  std::vector<estimatorBase *> createEstimators(const json_t &j);

 private:
  std::map<std::string, waveCreator> waves;
  std::map<std::string, potCreator> pots;
  std::map<std::string, obCreator> obs;
};
factory &getFactory();
void registerDefaults(factory &f);      // everything dmc/main.cpp:188-243 registers that is on the DMC path

// ---- driver -------------------------------------------------------------------------------------------
struct seedHolder {
  void seed(unsigned long long s) { value = s; }
  unsigned long long value = 0;
};

class dmcDriver {
 public:
  dmcDriver(wavefunction_t *wave, potential_t *pot, real_t timeStep, size_t nWalkers, size_t delta_walkers);
  ~dmcDriver();
  size_t &getStepsPerBlock() { return _stepsPerBlock; }
  size_t &getCorrelationSteps() { return correlationSteps; }
  std::vector<estimatorBase *> &getEstimators() { return estimators; }
  seedHolder &getRandomGenerator() { return rng; }
  void disableBranching() { performBranching = false; }
  // method "vmc" (vmcDriver.cpp): nWalkers independent Metropolis chains with the Gaussian mover of vmcMoves.cpp:11-35
  // (sigma = timeStep) and the |psi|^2 acceptance; the block report is vmcDriver::out's (no Step / Curr. Walkers lines)
  void enableVmc() { vmcMode = true; performBranching = false; }
  void setDevice(int d) { device = d; }
  void setComm(const void *uniqueId, int rank_, int nranks_) { uid = uniqueId; rank = rank_; nranks = nranks_; }
  void setOutput(bool on) { master = on; }
  /* random initial conditions are redrawn until the wavefunction's constraints hold (main.cpp:253-266): attempt k = 0, 1, ... */
  void setInitialConditionGenerator(std::function<states_t(unsigned)> g) { redraw = std::move(g); }
  void run(const std::vector<states_t> &states, size_t nBlocks);
  double walkerMovesPerSecond() const { return moves / (deviceMs * 1e-3); }

 private:
  void check(int rc) const;
  void out(size_t iBlock);
  void dumpConfigurations() const;
  void writeWaveTables() const;
  wavefunction_t *psi;
  potential_t *pot;
  real_t timeStep;
  size_t nWalkers, deltaWalkers;
  size_t _stepsPerBlock = 0, correlationSteps = 0;
  std::vector<estimatorBase *> estimators;
  std::unique_ptr<realScalarEstimator> energyEst;
  seedHolder rng;
  bool performBranching = true, master = true, vmcMode = false;
  int device = 0, rank = 0, nranks = 1;
  const void *uid = nullptr;
  std::function<states_t(unsigned)> redraw;
  qmc_ctx *ctx = nullptr;
  int dofs = 0;
  std::vector<int> speciesN;
  double moves = 0, deviceMs = 0, totalMs = 0, outputMs = 0;
};

std::string ansiColor(const std::string &color_name);
std::vector<states_t> readStatesFromDirectory(const std::string &dirname, int rank, int nranks);
