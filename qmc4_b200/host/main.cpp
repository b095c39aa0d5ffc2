// dmc_b200 — the host driver: `dmc_b200 < input.json` or `dmc_b200 input.json [--gpus N]`, the reference's
// CLI and input schema (dmc/main.cpp:105-428) for method "dmc" / "svmc" / "vmc", running on B200 through the C ABI.
// Multi-GPU: one host thread per GPU inside this process stands for one MPI rank of the reference
// (seed, population split and per-rank configuration files follow main.cpp:383,391 and walkers.cpp:225).
#include "qmc4host.h"

#include <cstring>
#include <random>
#include <sstream>
#include <thread>

static std::string readToString(std::istream &in)
{
  std::string s, line;
  while (std::getline(in, line)) s += line;
  return s;
}

static states_t generateRandomInitialConditions(const std::vector<int> &Ns, const std::vector<real_t> &lBoxInit, unsigned long long seed)
{ /* main.cpp:69-91: uniform in [-1,1) * lBox/2 per dimension (the reference draws from rand()) */
  std::mt19937_64 gen(seed);
  std::uniform_real_distribution<real_t> u(-1., 1.);
  states_t states;
  for (int n : Ns)
    {
      state_t s(3 * (size_t)n);
      for (int i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d) s[(size_t)d * n + i] = u(gen) * lBoxInit[d] * 0.5;
      states.push_back(s);
    }
  return states;
}

static void runRank(const json_t &j, int rank, int nranks, const void *uid, int *status, std::string *err)
{
  try
    {
      std::vector<real_t> lBox = j["lBox"].get<std::vector<real_t> >();
      std::vector<int> Ns = j["N"].get<std::vector<int> >();
      std::vector<real_t> lBoxInit = lBox;
      if (j.contains("initialConditionGenerator") && j["initialConditionGenerator"].contains("lBox"))
        lBoxInit = j["initialConditionGenerator"]["lBox"].get<std::vector<real_t> >();
      if ((int)lBox.size() != getDimensions() || (int)lBoxInit.size() != getDimensions())
        throw invalidInput("Input file implies dimensionality different from " + std::to_string(getDimensions()));
      geometryPBC geo(lBox[0], lBox[1], lBox[2]);
      factory f = getFactory();                      // per-rank copy of the registered types
      auto waves = f.createWavefunctions(j["wavefunctions"], geo);
      productWavefunction psi(waves);
      auto potentials = f.createPotentials(j["potentials"], geo);
      sumPotentials pot(potentials);
      const std::string method = j["method"].get<std::string>();
      const real_t timeStep = j["timeStep"].get<real_t>();
      const size_t stepsPerBlock = j["stepsPerBlock"].get<int>();
      const size_t nBlocks = j["nBlocks"].get<int>();
      const int seed = j["seed"].get<int>();
      const size_t correlationSteps = j["correlationSteps"].get<int>();       /* required (main.cpp:287) */
      std::vector<states_t> configurations;
      if (j.contains("configurations"))
        {
          configurations = readStatesFromDirectory(j["configurations"].get<std::string>(), rank, nranks);
          if (configurations.empty() && j.contains("initialConfigurations"))
            configurations = readStatesFromDirectory(j["initialConfigurations"].get<std::string>(), rank, nranks);
          if (rank == 0 && configurations.empty())
            std::cout << ansiColor("yellow") << "WARNING: Failed to load configurations from directory. Falling back to randomly generated configurations."
                      << ansiColor("default") << std::endl;
        }
      if (method != "dmc" && method != "svmc" && method != "vmc")
        throw missingImplementation("method '" + method + "': dmc, svmc and vmc run on the GPU path");
      const bool vmc = method == "vmc";
      if (vmc && j.contains("measurements"))
        for (const json_t &m : j["measurements"].elements())
          if (m["kind"].get<std::string>() == "superfluidFraction")
            throw invalidInput("Cannot have vmc and superfluid fraction measurements");        /* main.cpp:279-282 */
      /* every rank registers the estimators (the device accumulates everywhere and the block reduction is a
         collective); only rank 0 writes the files (estimators.cpp:40,58) */
      std::vector<estimatorBase *> ests = f.createEstimators(j.contains("measurements") ? j["measurements"] : json_t::parse("[]"));
      /* vmc: the reference runs ONE chain per MPI rank (vmcDriver.cpp:21-36) and does not read "walkers"; on the GPU the
         optional "walkers" key asks for that many independent chains (extension; default one chain per GPU) */
      const size_t nW = vmc ? (size_t)(j.contains("walkers") ? j["walkers"].get<int>() : nranks) : (size_t)j["walkers"].get<int>();
      size_t delta_walkers = size_t(nW * 0.1);
      if (j.contains("delta_walkers")) delta_walkers = j["delta_walkers"].get<int>();
      dmcDriver dmcO(&psi, &pot, timeStep, nW, delta_walkers);
      dmcO.getStepsPerBlock() = stepsPerBlock;
      dmcO.getCorrelationSteps() = correlationSteps;
      dmcO.setDevice(rank);
      dmcO.setOutput(rank == 0);
      if (nranks > 1) dmcO.setComm(uid, rank, nranks);
      std::vector<states_t> dmcStates(configurations);
      if (dmcStates.empty())
        { /* all walkers start from the same random configuration (main.cpp:380-388) */
          states_t states = generateRandomInitialConditions(Ns, lBoxInit, (unsigned long long)seed);
          for (size_t i = 0; i < nW / nranks; i++) dmcStates.push_back(states);
          /* ... redrawn (same sequence on every rank) until the wavefunction's constraints are satisfied (main.cpp:253-266) */
          dmcO.setInitialConditionGenerator([Ns, lBoxInit, seed](unsigned attempt) {
            return generateRandomInitialConditions(Ns, lBoxInit, (unsigned long long)seed + 7919ull * attempt);
          });
        }
      for (const auto &st : dmcStates)
        for (size_t s = 0; s < Ns.size(); ++s)
          if (s >= st.size() || getN(st[s]) != Ns[s])
            throw invalidInput("Initial configuration does not math the numper of particles defined in the input file");
      dmcO.getRandomGenerator().seed((unsigned long long)seed + rank);
      if (method == "svmc") dmcO.disableBranching();
      if (vmc) dmcO.enableVmc();
      /* main.cpp:270-273,318-335,395-408: the forceEnergy estimator always exists (its file gets the header), it is attached
         for vmc and svmc only, before the estimators of "measurements" */
      std::unique_ptr<forceEnergyEstimator> m2(new forceEnergyEstimator);
      if (rank == 0) m2->writeHeaderNow();
      bool wantForceEnergy = false;
      if (j.contains("measurements"))
        for (const json_t &m : j["measurements"].elements())
          if (m["kind"].get<std::string>() == "forceEnergy") wantForceEnergy = true;
      if (wantForceEnergy && (vmc || method == "svmc")) dmcO.getEstimators().push_back(m2.release());   /* the driver owns its estimators */
      for (auto *e : ests) dmcO.getEstimators().push_back(e);
      dmcO.run(dmcStates, nBlocks);
      for (auto *w : waves) delete w;
      for (auto *p : potentials) delete p;
      *status = 0;
    }
  catch (const std::exception &e)
    {
      *status = 1;
      *err = e.what();
    }
}

int main(int argc, char **argv)
{
  std::string text;
  int gpus = 1;
  const char *file = nullptr;
  for (int i = 1; i < argc; ++i)
    {
      if (!strcmp(argv[i], "--gpus") && i + 1 < argc) gpus = atoi(argv[++i]);
      else file = argv[i];
    }
  if (file)
    {
      std::ifstream fd(file);
      if (!fd)
        {
          std::cerr << "cannot open " << file << std::endl;
          return 2;
        }
      text = readToString(fd);
    }
  else text = readToString(std::cin);
  try
    {
      const json_t j = json_t::parse(text);
      std::cout << "MPI processes: " << gpus << std::endl;      /* main.cpp:146-149 (ranks are GPUs here) */
      registerDefaults(getFactory());
      unsigned char uid[128] = {0};
      if (gpus > 1 && qmc_comm_unique_id(uid)) throw invalidState(qmc_last_error(nullptr));
      std::vector<std::thread> th;
      std::vector<int> status(gpus, 0);
      std::vector<std::string> errs(gpus);
      for (int r = 1; r < gpus; ++r) th.emplace_back(runRank, std::cref(j), r, gpus, (const void *)uid, &status[r], &errs[r]);
      runRank(j, 0, gpus, uid, &status[0], &errs[0]);
      for (auto &t : th) t.join();
      for (int r = 0; r < gpus; ++r)
        if (status[r])
          {
            std::cerr << "terminate called after throwing an exception on rank " << r << ": " << errs[r] << std::endl;
            return 1;
          }
    }
  catch (const std::exception &e)
    {
      std::cerr << "terminate called after throwing an exception: " << e.what() << std::endl;
      return 1;
    }
  return 0;
}
