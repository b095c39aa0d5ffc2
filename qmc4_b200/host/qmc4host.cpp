#include "qmc4host.h"

#include <dirent.h>
#include <sys/stat.h>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <iomanip>
#include <regex>
#include <sstream>

// ---- ids ------------------------------------------------------------------------------------------------
std::string createId(const json_t &j)
{ /* dmc/abstractFactory.cpp:3-20 */
  std::string id = "";
  if (j.is_object() && j.contains("kind")) id += j["kind"].get<std::string>();
  for (const json_t *element : j.children())
    if (element->is_structured())
      {
        const std::string id2 = createId(*element);
        if (id2 != "") id += "/" + id2;
      }
  return id;
}

static void rethrow(qmc_ctx *ctx, int rc)
{
  if (!rc) return;
  const std::string m = qmc_last_error(ctx);
  if (m.find("Inalid initial") != std::string::npos || m.find("two body jastrow") != std::string::npos) throw invalidInput(m);
  if (m.find("unknown") != std::string::npos) throw missingImplementation(m);
  throw invalidState(m);
}

void wavefunction::add(qmc_ctx *ctx, int wfKind, const std::vector<int> &s, const jastrowBase &J, real_t hs) const
{
  rethrow(ctx, qmc_add_wavefunction_term(ctx, wfKind, s.data(), (int)s.size(), J.kind, J.params.data(), (int)J.params.size(), hs));
}
void potential::add(qmc_ctx *ctx, int kind, const std::vector<int> &s, const std::vector<real_t> &p) const
{
  rethrow(ctx, qmc_add_potential_term(ctx, kind, s.data(), (int)s.size(), p.data(), (int)p.size()));
}
int pairCorrelation::emit(qmc_ctx *ctx, int bins, real_t minx, real_t maxx) const
{
  int idx = -1;
  const int s[2] = {setA, setB};
  rethrow(ctx, qmc_add_estimator(ctx, QMC_EST_PAIRCORRELATION, s, 2, bins, minx, maxx, nullptr, 0, &idx));
  return idx;
}
int trimerCorrelationUnDis::emit(qmc_ctx *ctx, int bins, real_t minx, real_t maxx) const
{
  int idx = -1;
  rethrow(ctx, qmc_add_estimator(ctx, QMC_EST_TRIMERCORRELATION_UNDIS, &setA, 1, bins, minx, maxx, nullptr, 0, &idx));
  return idx;
}
template <int EST_KIND, bool SCALAR> int threeSetObservable<EST_KIND, SCALAR>::emit(qmc_ctx *ctx, int bins, real_t minx, real_t maxx) const
{
  int idx = -1;
  const int s[3] = {setA, setB, setC};
  rethrow(ctx, qmc_add_estimator(ctx, EST_KIND, s, 3, SCALAR ? 1 : bins, minx, maxx, nullptr, 0, &idx));
  return idx;
}
template class threeSetObservable<QMC_EST_TRIMERCORRELATION, false>;
template class threeSetObservable<QMC_EST_TRIMERCORRELATION_DIS, false>;
template class threeSetObservable<QMC_EST_TRIMERHYPERRADIUS, true>;
int centerOfMassSquared::emit(qmc_ctx *ctx, int, real_t, real_t) const
{
  int idx = -1;
  rethrow(ctx, qmc_add_estimator(ctx, QMC_EST_CENTEROFMASSSQUARED, &setA, 1, 1, 0, 1, nullptr, 0, &idx));
  return idx;
}
std::vector<real_t> structureFactor::x(int n) const
{ /* build_q_vector (structureFactor.cpp:8-40) */
  double dq = int(maxq / (2 * M_PI * n / lBox)) * 2 * M_PI / lBox;
  if (dq == 0) dq = 2 * M_PI / lBox;
  std::vector<real_t> q(n);
  for (int i = 0; i < n; ++i) q[i] = 2 * M_PI / lBox + (i == 0 ? 0 : i * dq);
  return q;
}
int structureFactor::emit(qmc_ctx *ctx, int nb, real_t, real_t) const
{
  int idx = -1;
  const int s[2] = {setA, setB};
  const double extra[3] = {maxq, lBox, (double)channel};
  rethrow(ctx, qmc_add_estimator(ctx, QMC_EST_STRUCTUREFACTOR, s, 2, nb, 0, 0, extra, 3, &idx));
  return idx;
}

// ---- estimators (dmc/estimators.cpp) ------------------------------------------------------------------
static bool file_is_empty(const std::string &name)
{
  std::ifstream of(name);
  return of.peek() == std::ifstream::traits_type::eof();
}
estimatorBase::estimatorBase(const std::string &label) : _label(label), filename(label + ".dat")
{
  f.open(filename, std::fstream::out | std::fstream::app);
}
void estimatorBase::headerIfEmpty(const std::string &h)
{ /* the reference writes the header in the constructor on rank 0 only (estimators.cpp:35-46); here it is
     deferred to the first dump(), which only the master rank performs */
  if (file_is_empty(filename)) pendingHeader = h;
}
void estimatorBase::dump()
{
  if (!pendingHeader.empty())
    {
      f << pendingHeader << std::endl;
      pendingHeader.clear();
    }
  write(f);
  f << std::endl;
}
realScalarEstimator::realScalarEstimator(const std::string &label) : estimatorBase(label) { headerIfEmpty(label); }
void realScalarEstimator::write(std::ostream &stream)
{
  stream << std::scientific << std::setprecision(12);
  if (n > 0) stream << sum / n;
}
realScalarObservableEstimator::realScalarObservableEstimator(observableBase *ob_, const json_t &j)
    : estimatorBase(j["label"].get<std::string>()), ob(ob_)
{
  headerIfEmpty(_label);                               /* estimators.cpp:35-46 */
}
void realScalarObservableEstimator::collect(qmc_ctx *ctx, const qmc_block_result &) { rethrow(ctx, qmc_get_estimator(ctx, index, &sum, 1, &n)); }
void realScalarObservableEstimator::write(std::ostream &stream)
{ /* realScalarEstimator::write (estimators.cpp:89-98) */
  stream << std::scientific << std::setprecision(12);
  if (n > 0) stream << sum / n;
}
realHistogramEstimator::realHistogramEstimator(observableBase *ob_, const json_t &j)
    : estimatorBase(j["label"].get<std::string>()), ob(ob_), vectorStyle(ob_->isVector())
{
  const size_t size = j["bins"].get<size_t>();
  minx = vectorStyle ? 0 : j["minx"].get<real_t>();
  maxx = vectorStyle ? (real_t)size : j["maxx"].get<real_t>();
  sums.assign(size, 0);
  headerIfEmpty("x " + _label);
  if (vectorStyle) x = ob->x((int)size);
  else
    {
      const real_t step = (maxx - minx) / size;
      x.resize(size);
      for (size_t i = 0; i < size; i++) x[i] = step * i;      /* left edge, ignores minx (estimators.cpp:66-70) */
    }
}
void realHistogramEstimator::collect(qmc_ctx *ctx, const qmc_block_result &)
{
  rethrow(ctx, qmc_get_estimator(ctx, index, sums.data(), (int)sums.size(), &weight));
}
void realHistogramEstimator::write(std::ostream &stream)
{
  if (vectorStyle) stream << std::scientific << std::setprecision(12);     /* realVectorEstimator::write (estimators.cpp:101-110) */
  for (size_t i = 0; i < x.size(); i++) stream << x[i] << " " << sums[i] / weight << std::endl;
}

void forwardWalkingGroup::ensureRegistered(qmc_ctx *ctx)
{
  if (first >= 0) return;
  const int target = ob->emit(ctx, bins, minx, maxx);
  rethrow(ctx, qmc_add_forward_walking(ctx, target, recordSteps, steps.data(), (int)steps.size(), &first));
}
realHistogramForwardWalkingEstimator::realHistogramForwardWalkingEstimator(std::shared_ptr<forwardWalkingGroup> g, int k_,
                                                                           const std::string &label)
    : estimatorBase(label), group(g), k(k_)
{
  sums.assign(group->bins, 0);
  headerIfEmpty("x " + _label);                       /* correlationEstimator.cpp:160-168 */
}
void realHistogramForwardWalkingEstimator::collect(qmc_ctx *ctx, const qmc_block_result &)
{
  rethrow(ctx, qmc_get_estimator(ctx, group->first + k, sums.data(), (int)sums.size(), &weight));
}
void realHistogramForwardWalkingEstimator::write(std::ostream &stream)
{ /* correlationEstimator.cpp:171-178 -> histogramAccumulator::writeAverage (accumulators.h:146-156): bin CENTRES */
  if (!(weight > 0)) return;
  const real_t dx = (group->maxx - group->minx) / group->bins;
  for (size_t i = 0; i < sums.size(); i++) stream << group->minx + (i + 0.5) * dx << " " << sums[i] / weight << std::endl;
}

realScalarForwardWalkingEstimator::realScalarForwardWalkingEstimator(std::shared_ptr<forwardWalkingGroup> g, int k_, const std::string &label)
    : estimatorBase(label), group(g), k(k_)
{
  headerIfEmpty(_label);                               /* correlationEstimator.cpp:53-66 */
}
void realScalarForwardWalkingEstimator::collect(qmc_ctx *ctx, const qmc_block_result &) { rethrow(ctx, qmc_get_estimator(ctx, group->first + k, &sum, 1, &n)); }
void realScalarForwardWalkingEstimator::write(std::ostream &stream)
{ /* correlationEstimator.cpp:73-80: default stream formatting (no std::scientific here in the reference) */
  if (n > 0) stream << sum / n;
}

forceEnergyEstimator::forceEnergyEstimator() : estimatorBase("forceEnergy") { headerIfEmpty(_label); }
void forceEnergyEstimator::writeHeaderNow()
{
  if (pendingHeader.empty()) return;
  f << pendingHeader << std::endl;
  pendingHeader.clear();
}
void forceEnergyEstimator::registerOn(qmc_ctx *ctx) { rethrow(ctx, qmc_enable_force_energy(ctx)); }
void forceEnergyEstimator::collect(qmc_ctx *ctx, const qmc_block_result &) { rethrow(ctx, qmc_get_force_energy(ctx, &sum, &n)); }
void forceEnergyEstimator::write(std::ostream &stream)
{ /* realScalarEstimator::write (estimators.cpp:89-98) */
  stream << std::scientific << std::setprecision(12);
  if (n > 0) stream << sum / n;
}

superfluidFractionEstimator::superfluidFractionEstimator(const json_t &j) : estimatorBase(j["label"].get<std::string>())
{
  const json_t &st = j["sets"];
  if (st.size() != 2) throw invalidInput("Superfluid fraction should have two sets");      /* estimatorsFactory.cpp:88-91 */
  setA = st[0].get<int>();
  setB = st[1].get<int>();
  const size_t bins = j["bins"].get<size_t>();
  sums.assign(bins, 0);
  weights.assign(bins, 0);
  headerIfEmpty("x " + _label);                       /* correlationEstimator.cpp:280-289 */
}
void superfluidFractionEstimator::registerOn(qmc_ctx *ctx)
{
  rethrow(ctx, qmc_add_superfluid_fraction(ctx, setA, setB, (int)sums.size(), &index));
}
void superfluidFractionEstimator::collect(qmc_ctx *ctx, const qmc_block_result &)
{
  rethrow(ctx, qmc_get_vector_estimator(ctx, index, sums.data(), weights.data(), (int)sums.size()));
}
void superfluidFractionEstimator::write(std::ostream &stream)
{ /* correlationEstimator.cpp:305-320: lag index and average of the bins that received data, default formatting */
  for (size_t i = 0; i < sums.size(); i++)
    if (weights[i] > 0) stream << i << " " << sums[i] / weights[i] << std::endl;
}

// ---- factory ---------------------------------------------------------------------------------------------
factory &getFactory()
{
  static factory singleton;
  return singleton;
}
void registerDefaults(factory &f)
{
  f.registerJastrow<gaussianJastrow>();
  f.registerJastrow<jastrowGaussGauss>();
  f.registerJastrow<jastrowSquareWell>();
  f.registerJastrow<jastrowSquareWellBoundState>();
  f.registerJastrow<jastrowPoschTeller>();
  f.registerJastrow<jastrowHardSphereGauss>();
  f.registerJastrow<jastrowHardSphere>();
  f.registerJastrow<jastrowSpline>();
  f.registerJastrow<jastrow_delta_phonons>();                /* main.cpp:213-215 */
  f.registerJastrow<jastrow_delta_in_trap>();
  f.registerJastrow<jastrow_delta_bound_state_phonons>();
  f.registerJastrow<jastrowDipolarRep>();                    /* main.cpp:216 */
  f.registerObservable<pairCorrelation>();
  f.registerObservable<trimerCorrelationUnDis>();
  f.registerObservable<structureFactor>();
  f.registerObservable<centerOfMassSquared>();               /* main.cpp:203-207 */
  f.registerObservable<trimerhyperRadius>();
  f.registerObservable<trimerCorrelation>();
  f.registerObservable<trimerCorrelationDis>();
  f.registerPotential<harmonicPotential>();
  f.registerPotential<squareWellPotential2b>();
  f.registerPotential<dipolarPotential2b>();
  f.registerPotential<poschTellerPotential2b>();
  f.registerPotentialFunctor<squareWell>();
  f.registerPotentialFunctor<potentialBarrier>();
}
std::vector<wavefunction *> factory::createWavefunctions(const json_t &j, const geometry_t &geo)
{
  std::vector<wavefunction *> out;
  for (const json_t &w : j.elements())
    {
      const std::string id = createId(w);
      auto it = waves.find(id);
      if (it == waves.end()) throw factoryIdNotRecorded(id);
      out.push_back(it->second(w, geo));
    }
  return out;
}
std::vector<potential *> factory::createPotentials(const json_t &j, const geometry_t &geo)
{
  std::vector<potential *> out;
  for (const json_t &p : j.elements())
    {
      const std::string id = createId(p);
      auto it = pots.find(id);
      if (it == pots.end()) throw factoryIdNotRecorded(id);
      out.push_back(it->second(p, geo));
    }
  return out;
}
std::vector<estimatorBase *> factory::createEstimators(const json_t &j)
{ /* estimatorFactory::create (estimatorsFactory.cpp:39-133): keyed by plain "kind"; energy/forceEnergy skipped */
  std::vector<estimatorBase *> out;
  for (const json_t &e : j.elements())
    {
      const std::string id = e["kind"].get<std::string>();
      if (e.contains("forwardWalkingSteps"))
        { /* estimatorsFactory.cpp:47-74 */
          auto it = obs.find(id);
          if (it == obs.end()) throw factoryIdNotRecorded(id);
          auto g = std::make_shared<forwardWalkingGroup>();
          g->ob.reset(it->second(e));
          if (g->ob->isVector()) throw missingImplementation("Forward walking estimator for vector not yet implemented");
          if (!e.contains("recordSteps") || e["recordSteps"].get<int>() <= 0)
            throw invalidInput("forwardWalkingSteps needs recordSteps > 0 on the same measurement (the storer, estimatorsFactory.h:112-121)");
          const bool scalar = g->ob->isScalar();
          g->bins = scalar ? 1 : e["bins"].get<int>();
          g->minx = scalar ? 0 : e["minx"].get<real_t>();
          g->maxx = scalar ? 1 : e["maxx"].get<real_t>();
          g->recordSteps = e["recordSteps"].get<int>();
          for (const json_t &st : e["forwardWalkingSteps"].elements()) g->steps.push_back(st.get<int>());
          const std::string label = e["label"].get<std::string>();
          for (size_t k = 0; k < g->steps.size(); k++)
            {
              const std::string lk = label + "_fw" + std::to_string(g->steps[k]);
              if (scalar) out.push_back(new realScalarForwardWalkingEstimator(g, (int)k, lk));      /* estimatorsFactory.cpp:60-63 */
              else out.push_back(new realHistogramForwardWalkingEstimator(g, (int)k, lk));
            }
          continue;
        }
      if (id == "superfluidFraction")
        { /* estimatorsFactory.cpp:80-116 */
          out.push_back(new superfluidFractionEstimator(e));
          continue;
        }
      if (id == "forceEnergy" || id == "energy") continue;
      auto it = obs.find(id);
      if (it == obs.end()) throw factoryIdNotRecorded(id);
      observableBase *ob = it->second(e);
      if (ob->isScalar()) out.push_back(new realScalarObservableEstimator(ob, e));
      else out.push_back(new realHistogramEstimator(ob, e));
    }
  return out;
}

// ---- tools (dmc/tools.cpp) ----------------------------------------------------------------------------
std::string ansiColor(const std::string &c)
{
  if (c == "red") return "\033[0;31m";
  if (c == "default") return "\033[0m";
  if (c == "cyan") return "\033[0;36m";
  if (c == "green") return "\033[0;32m";
  if (c == "yellow") return "\033[33m";
  throw invalidInput(c);
}

static std::vector<states_t> readStates(const json_t &jI)
{ /* tools.cpp:55-96: {"configurations":[ [ {"x":[],"y":[],"z":[]} per species ] per walker ]} */
  std::vector<states_t> states;
  for (const json_t &jState : jI.elements())
    {
      states_t st;
      for (const json_t &jSet : jState.elements())
        {
          const auto x = jSet["x"].get<std::vector<double> >(), y = jSet["y"].get<std::vector<double> >(), z = jSet["z"].get<std::vector<double> >();
          state_t s;
          s.insert(s.end(), x.begin(), x.end());
          s.insert(s.end(), y.begin(), y.end());
          s.insert(s.end(), z.begin(), z.end());
          st.push_back(s);
        }
      states.push_back(st);
    }
  return states;
}

std::vector<states_t> readStatesFromDirectory(const std::string &dirname, int rank, int nranks)
{ /* tools.cpp:98-161: files matching .*-Rank([0-9]+).* are redistributed over the ranks */
  std::vector<states_t> states;
  DIR *d = opendir(dirname.c_str());
  if (!d) return states;
  std::vector<json_t> all;
  std::regex rgx(".*-Rank([0-9]+).*");
  std::vector<std::string> names;
  while (dirent *e = readdir(d)) names.push_back(e->d_name);
  closedir(d);
  std::sort(names.begin(), names.end());
  for (const auto &n : names)
    if (std::regex_match(n, rgx)) all.push_back(json_t::loadFromFile(dirname + "/" + n));
  if (all.empty()) return states;
  const size_t old = all.size();
  for (size_t i = 0; all.size() < (size_t)nranks; ++i) all.push_back(all[i % old]);
  const size_t nLocal = all.size() / nranks;
  auto take = [&](size_t i) {
    auto loc = readStates(all[i]["configurations"]);
    states.insert(states.end(), loc.begin(), loc.end());
  };
  for (size_t i = nLocal * rank; i < nLocal * (rank + 1); ++i) take(i);
  if ((size_t)rank < all.size() % nranks) take(nLocal * nranks + rank);
  return states;
}

// ---- dmcDriver (dmc/dmcDriver.cpp, dmc/driver.cpp) -------------------------------------------------------
dmcDriver::dmcDriver(wavefunction_t *wave, potential_t *pot_, real_t timeStep_, size_t nW, size_t delta)
    : psi(wave), pot(pot_), timeStep(timeStep_), nWalkers(nW), deltaWalkers(delta)
{
}
dmcDriver::~dmcDriver()
{
  if (ctx) qmc_destroy(ctx);
  for (auto *e : estimators) delete e;
}
void dmcDriver::check(int rc) const { rethrow(ctx, rc); }

void dmcDriver::run(const std::vector<states_t> &states, size_t nBlocks)
{
  if (states.empty()) throw invalidInput("no initial configurations");
  const geometry_t &geo = psi->getGeometry();
  const double lBox[3] = {geo.getLBox(0), geo.getLBox(1), geo.getLBox(2)};
  std::vector<int> N;
  for (const auto &s : states[0]) N.push_back(getN(s));
  if (qmc_create(&ctx, device, lBox, (int)N.size(), N.data())) throw invalidState(qmc_last_error(nullptr));
  for (const wavefunction *w : psi->waves()) w->emit(ctx);
  for (const potential *p : pot->potentials()) p->emit(ctx);
  if (master) energyEst.reset(new realScalarEstimator("energy"));        /* registered first (dmcDriver.cpp:245) */
  for (auto *e : estimators) e->registerOn(ctx);
  if (nranks > 1) check(qmc_comm_init(ctx, uid, rank, nranks));
  if (vmcMode) check(qmc_set_vmc(ctx, timeStep, (int64_t)nWalkers, rng.value, 0));
  else check(qmc_set_dmc(ctx, timeStep, (int64_t)nWalkers, (int64_t)deltaWalkers, rng.value, performBranching ? 1 : 0, 0));
  dofs = 0;
  speciesN = N;
  for (int n : N) dofs += 3 * n;
  std::vector<double> R(states.size() * (size_t)dofs);
  auto flatten = [&](const states_t &st, size_t w) {
    size_t off = 0;
    for (const auto &s : st)
      {
        std::copy(s.begin(), s.end(), R.begin() + w * dofs + off);
        off += s.size();
      }
  };
  for (size_t w = 0; w < states.size(); ++w) flatten(states[w], w);
  if (redraw)
    { /* all walkers start from ONE random configuration (main.cpp:380-388), redrawn until psi.satisfyConstraints holds
         (main.cpp:253-266); the constraint is evaluated by the device (qmc_eval's valid flag) */
      for (unsigned attempt = 0;; ++attempt)
        {
          int valid = 0;
          check(qmc_eval(ctx, R.data(), 1, nullptr, nullptr, nullptr, nullptr, nullptr, &valid));
          if (valid) break;
          if (attempt >= 100000) throw invalidState("no random initial configuration satisfies the constraints of the wavefunction");
          const states_t st = redraw(attempt + 1);
          for (size_t w = 0; w < states.size(); ++w) flatten(st, w);
        }
    }
  if (master) writeWaveTables();
  check(qmc_upload_walkers(ctx, R.data(), (int64_t)states.size()));      /* initialize + setEnergyShift (dmcDriver.cpp:163-191) */
  const auto t0 = std::chrono::steady_clock::now();
  for (size_t iBlock = 0; iBlock < nBlocks; ++iBlock)
    { /* driver::run (driver.cpp:23-43): stepsPerBlock x (correlationSteps x step, then step + accumulate) */
      check(qmc_step(ctx, (int)(_stepsPerBlock * (correlationSteps + 1)), (int)(correlationSteps + 1)));
      const auto o0 = std::chrono::steady_clock::now();
      out(iBlock);
      outputMs += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - o0).count();
    }
  totalMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (master)
    { /* timer report (driver.cpp:45-57); the five timer names of the reference */
      auto line = [&](const char *n, double ms) { std::cout << n << " :  " << ms << " ms ( " << ms / totalMs * 100 << " %  )" << std::endl; };
      line("waitWalkers", 0);
      line("sendWalkers", 0);
      line("shiftEnergy", 0);
      line("outputTime", outputMs);
      line("totalRunningTime", totalMs);
      std::cout << "deviceTime :  " << deviceMs << " ms ; walker-moves/s : " << moves / (deviceMs * 1e-3) << std::endl;
    }
}

void dmcDriver::out(size_t iBlock)
{ /* dmcDriver::out (dmcDriver.cpp:197-230) */
  qmc_block_result r;
  check(qmc_block_results(ctx, &r));
  moves += r.n_moves;
  deviceMs += r.device_ms;
  if (r.overflow != 0) throw invalidState("walker population exceeded the device buffer capacity");
  for (auto *e : estimators) e->collect(ctx, r);
  int64_t moved = 0;
  /* walkers migrate inside qmc_step, one step late (dmcDriver.cpp:112,150-157); this synchronous pass at the block boundary
     evens out what the dead band left and is the path that carries forward-walking records along */
  if (nranks > 1) check(qmc_rebalance(ctx, &moved));
  if (master)
    {
      energyEst->collect(ctx, r);
      std::cout << ansiColor("green") << "Block: " << ansiColor("default") << iBlock << std::endl;
      std::cout << "Acc. Ratio: " << r.n_accepted / r.n_moves << std::endl;
      if (!vmcMode) std::cout << "Step: " << (iBlock + 1) * _stepsPerBlock << std::endl;      /* vmcDriver::out (vmcDriver.cpp:96-121) has no Step line */
      std::cout << ansiColor("cyan") << "Energy: " << ansiColor("default");
      std::cout << std::scientific;
      energyEst->write(std::cout);
      std::cout << std::endl << std::defaultfloat;
      if (!vmcMode)
        {
          std::cout << ansiColor("cyan") << "Curr. Walkers: " << ansiColor("default");
          double globalPopulation = r.population;        /* the reference prints the sum over the ranks (dmcDriver.cpp:204,222) */
          if (nranks > 1) check(qmc_get_info(ctx, "global_population", &globalPopulation));
          std::cout << (long)globalPopulation << std::endl;
        }
      energyEst->dump();
      for (auto *e : estimators) e->dump();
    }
  check(qmc_clear_estimators(ctx));
  dumpConfigurations();
}

void dmcDriver::writeWaveTables() const
{ /* main.cpp:220-227: wave<i>.dat = waves[i]->print(), the radial functor tabulated at the bin centres of [0, lBox/2]
     (jastrow.h:42-63) — here by the device, so the file shows what the kernels evaluate */
  const real_t half = psi->getGeometry().getLBox(0) / 2.;
  std::vector<double> tab;
  for (size_t i = 0; i < psi->waves().size(); ++i)
    {
      std::ofstream f("wave" + std::to_string(i) + ".dat");
      const size_t n = psi->waves()[i]->printPoints();
      if (n == 0) continue;
      tab.resize(4 * n);
      check(qmc_tabulate_wavefunction(ctx, (int)i, 0., half, (int64_t)n, tab.data()));
      for (size_t k = 0; k < n; ++k) f << tab[4 * k] << " " << tab[4 * k + 1] << " " << tab[4 * k + 2] << " " << tab[4 * k + 3] << std::endl;
    }
}

void dmcDriver::dumpConfigurations() const
{ /* walkerContainer::dump (walkers.cpp:216-250): configurations/walkers-Rank<r>.json, overwritten every block */
  int64_t n = 0;
  rethrow(ctx, qmc_download_walkers(ctx, nullptr, 0, &n, nullptr));
  std::vector<double> R((size_t)n * dofs);
  if (n) rethrow(ctx, qmc_download_walkers(ctx, R.data(), n, &n, nullptr));
  struct stat st;
  if (stat("configurations", &st) == -1) mkdir("configurations", 0700);
  std::ofstream f("configurations/walkers-Rank" + std::to_string(rank) + ".json");
  f << std::setprecision(17) << "{\"configurations\":[";
  static const char *axes[3] = {"x", "y", "z"};
  for (int64_t w = 0; w < n; ++w)
    {
      f << (w ? "," : "") << "[";
      size_t off = (size_t)w * dofs;
      bool firstSet = true;
      for (int Nset : speciesN)
        {
          f << (firstSet ? "" : ",") << "{";
          for (int d = 0; d < 3; ++d)
            {
              f << (d ? "," : "") << "\"" << axes[d] << "\":[";
              for (int i = 0; i < Nset; ++i) f << (i ? "," : "") << R[off + (size_t)d * Nset + i];
              f << "]";
            }
          f << "}";
          off += 3 * (size_t)Nset;
          firstSet = false;
        }
      f << "]";
    }
  f << "]}";
}
