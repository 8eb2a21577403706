// host_mirror_test.cpp -- the reference's own GP tests (GPSpec.scala:13-76, KernelSpec.scala:34-58) and a
// seeded synthetic problem, written against the C++ host mirror (include/dynaml_b200.hpp) the way the
// Scala specs are written against dynaml-core.  Prints one JSON object; tests/test_gpu_cpp_mirror.py checks
// it against the oracle.
//   usage: host_mirror_test problem.bin     (int64 N, d, M; then X[N*d], y[N], Xs[M*d], all float64)
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "dynaml_b200.hpp"

using namespace dynaml;

static void print_vec(const char *name, const std::vector<double> &v, bool comma = true) {
  std::printf("\"%s\": [", name);
  for (size_t i = 0; i < v.size(); ++i) std::printf("%s%.17g", i ? ", " : "", v[i]);
  std::printf("]%s\n", comma ? "," : "");
}

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  std::FILE *f = std::fopen(argv[1], "rb");
  if (!f) return 3;
  int64_t hdr[3];
  if (std::fread(hdr, sizeof(int64_t), 3, f) != 3) return 4;
  const int64_t N = hdr[0], d = hdr[1], M = hdr[2];
  std::vector<double> X(N * d), y(N), Xs(M * d);
  if (std::fread(X.data(), 8, X.size(), f) != X.size() || std::fread(y.data(), 8, y.size(), f) != y.size() ||
      std::fread(Xs.data(), 8, Xs.size(), f) != Xs.size())
    return 5;
  std::fclose(f);

  std::printf("{\n");
  try {
    // ---- GPSpec "compute likelihood correctly": Dirac(0.5) + Dirac(0.5), mean v(0), 2 points -> log 2 pi
    {
      GPRegression::Data data = {{{0.0}, 0.0}, {{1.0}, 1.0}};
      GPRegression gp(std::make_shared<DiracKernel>(0.5), std::make_shared<DiracKernel>(0.5), data,
                      [](const Vec &v) { return v[0]; });
      std::printf("\"gpspec_likelihood\": %.17g,\n", gp.energy(gp._current_state()));
    }
    // ---- GPSpec "compute posterior pdf and error bars correctly": one training point, SE(1,1), noise 0
    {
      GPRegression::Data data = {{{1.0}, 1.0}};
      auto noise = std::make_shared<DiracKernel>(0.0);
      noise->block_all_hyper_parameters();
      GPRegression gp(std::make_shared<SEKernel>(1.0, 1.0), noise, data);
      MultGaussianPRV post = gp.predictiveDistribution({{2.0}});
      auto bars = gp.predictionWithErrorBars({{2.0}}, 1);
      std::printf("\"gpspec_pmean\": %.17g, \"gpspec_psigma\": %.17g, \"gpspec_lo\": %.17g, \"gpspec_hi\": %.17g,\n",
                  post.mu[0], post.covariance(0, 0), std::get<2>(bars[0]), std::get<3>(bars[0]));
    }
    // ---- KernelSpec: SE(1,1)(1, 0)
    std::printf("\"kernelspec_se\": %.17g,\n", SEKernel(1.0, 1.0).evaluate({1.0}, {0.0}));
    std::printf("\"kernelspec_laplace\": %.17g, \"kernelspec_cauchy\": %.17g,\n",
                LaplacianKernel(1.0).evaluate({1.0}, {0.0}), CauchyKernel(1.0).evaluate({1.0}, {1.5}));
    // ---- KernelSpec.scala:205-229: partitioned builder, eval2 = 1 / (1 + (x - y)^2) = CauchyKernel(1) on {0, 1}, 1 per block
    {
      CauchyKernel ck(1.0);
      ck.setBlockSizes(1, 1);
      const PartitionedMatrix k4 = ck.buildBlockedKernelMatrix({{0.0}, {1.0}});
      bool ok = k4.rows == 2 && k4.cols == 2 && k4.rowBlocks == 2 && k4.colBlocks == 2 && k4.blocks.size() == 3;
      for (const auto &kv : k4.blocks) ok = ok && kv.second(0, 0) == (kv.first.first == kv.first.second ? 1.0 : 0.5);
      const DenseMatrix dense = k4.toDenseMatrix();
      ok = ok && dense(0, 1) == 0.5 && dense(1, 0) == 0.5 && dense(1, 1) == 1.0;
      std::printf("\"kernelspec_partitioned_ok\": %s,\n", ok ? "true" : "false");
      // SVMKernel.scala:107-145: d/dsigma of the Cauchy kernel = 2 k^2 r^2 / sigma^3 = 0.5 at r = 1, sigma = 1
      const auto gm = ck.buildKernelGradMatrix({{0.0}, {1.0}});
      std::printf("\"kernelspec_grad_sigma_01\": %.17g, \"kernelspec_grad_sigma_00\": %.17g,\n", gm.at("sigma")(0, 1),
                  gm.at("sigma")(0, 0));
    }
    // ---- block / unblock plumbing (KernelSpec.scala:12-32)
    {
      SEKernel se(1.0, 1.0);
      se.block({se.hyper_parameters[0]});
      const bool a = se.effective_hyper_parameters()[0] == se.hyper_parameters.back();
      se.block_all_hyper_parameters();
      const bool b = se.effective_hyper_parameters().empty();
      se.block({});
      const bool c = se.effective_hyper_parameters().size() == 2;
      std::printf("\"block_unblock_ok\": %s,\n", (a && b && c) ? "true" : "false");
    }
    // ---- seeded synthetic problem: Matern-5/2 + Dirac noise
    GPRegression::Data data;
    std::vector<Vec> test;
    for (int64_t i = 0; i < N; ++i) data.push_back({Vec(X.begin() + i * d, X.begin() + (i + 1) * d), y[i]});
    for (int64_t i = 0; i < M; ++i) test.push_back(Vec(Xs.begin() + i * d, Xs.begin() + (i + 1) * d));
    GPRegression gp(std::make_shared<GenericMaternKernel>(2.0, 2), std::make_shared<DiracKernel>(0.3), data);
    Config h = gp._current_state();
    std::printf("\"energy\": %.17g,\n", gp.energy(h));
    Config g = gp.gradEnergy(h);
    std::printf("\"grad_l\": %.17g, \"grad_noiseLevel\": %.17g, \"grad_keys\": %zu,\n", g.at("l"), g.at("noiseLevel"),
                g.size());
    MultGaussianPRV post = gp.predictiveDistribution(test);
    print_vec("mu", post.mu);
    print_vec("cov", post.covariance.data);
    auto ci = post.confidenceInterval(2.0);
    print_vec("lo", ci.first);
    print_vec("hi", ci.second);
    // ---- GridSearch over (l, noiseLevel), then persist + predict from the persisted model
    GridSearch gs(gp);
    gs.setGridSize(3).setStepSize(0.2);
    Config init = {{"l", 2.6}, {"noiseLevel", 0.9}};
    auto land = gs.getEnergyLandscape(init);
    std::vector<double> energies;
    for (auto &kv : land) energies.push_back(kv.first);
    print_vec("landscape", energies);
    Config best = gs.optimize(init, {{"persist", "true"}}).second;
    std::printf("\"best_l\": %.17g, \"best_noiseLevel\": %.17g,\n", best.at("l"), best.at("noiseLevel"));
    std::printf("\"persisted_predict0\": %.17g,\n", gp.predict(test[0]));
    // ---- CSA and ML-II run and stay positive
    CoupledSimulatedAnnealing csa(gp, 42);
    csa.setMaxIterations(3).setGridSize(2).setStepSize(0.2);
    Config cbest = csa.optimize(init).second;
    std::printf("\"csa_energy\": %.17g,\n", gp.energy({{"p", 2.0}, {"l", cbest.at("l")}, {"noiseLevel", cbest.at("noiseLevel")}}));
    // ---- non-PD: NaN hyper-parameter -> +Inf energy, NaN gradient (AbstractGPRegressionModel.scala:243-251, 531-533)
    const double bad = gp.energy({{"p", 2.0}, {"l", std::nan("")}, {"noiseLevel", 0.3}});
    Config gbad = gp.gradEnergy({{"p", 2.0}, {"l", std::nan("")}, {"noiseLevel", 0.3}});
    std::printf("\"nonpd_energy_is_inf\": %s, \"nonpd_grad_is_nan\": %s,\n", std::isinf(bad) ? "true" : "false",
                (std::isnan(gbad.at("l")) && std::isnan(gbad.at("noiseLevel"))) ? "true" : "false");
    // ---- kernel algebra (LocalScalarKernel.scala:53-91): (SE * Cauchy + Matern * 0.6) as a device composite
    {
      std::shared_ptr<LocalScalarKernel> se = std::make_shared<SEKernel>(1.4, 1.1);
      std::shared_ptr<LocalScalarKernel> ca = std::make_shared<CauchyKernel>(2.0);
      std::shared_ptr<LocalScalarKernel> ma = std::make_shared<GenericMaternKernel>(1.7, 2);
      std::shared_ptr<LocalScalarKernel> k = se * ca + ma * 0.6;
      GPRegression cgp(k, std::make_shared<DiracKernel>(0.3), data);
      Config ch = cgp._current_state();
      std::printf("\"composite_energy\": %.17g,\n", cgp.energy(ch));
      Config cg = cgp.gradEnergy(ch);
      std::vector<double> ordered;
      for (const auto &name : k->effective_hyper_parameters()) ordered.push_back(cg.at(name));
      ordered.push_back(cg.at("noiseLevel"));
      print_vec("composite_grad", ordered);
      MultGaussianPRV cpost = cgp.predictiveDistribution(test);
      print_vec("composite_mu", cpost.mu);
    }
    // ---- GPBasisFuncRegressionModel: trend 0.3 + 0.5 x0 - 0.2 x1 x2 with prior covariance diag(0.5, 0.2, 0.3) + 0.02
    {
      auto basis = [](const Vec &x) { return Vec{1.0, x[0], x[1] * x[2]}; };
      DenseMatrix covB;
      covB.rows = covB.cols = 3;
      covB.data = {0.52, 0.02, 0.02, 0.02, 0.22, 0.02, 0.02, 0.02, 0.32};
      GPBasisFuncRegression bgp(std::make_shared<SEKernel>(1.6, 1.1), std::make_shared<DiracKernel>(0.3), data, basis,
                                Vec{0.3, 0.5, -0.2}, covB);
      std::printf("\"basis_energy\": %.17g,\n", bgp.energy(bgp._current_state()));
      MultGaussianPRV bpost = bgp.predictiveDistribution(test);
      print_vec("basis_mu", bpost.mu);
    }
    // ---- unsupported / bad arguments surface as exceptions, never as a CPU fallback
    bool threw = false;
    try {
      SEKernel se(1.0, 1.0);
      se.setHyperParameters({{"bandwidth", 2.0}});
    } catch (const std::invalid_argument &) {
      threw = true;
    }
    std::printf("\"missing_key_throws\": %s\n", threw ? "true" : "false");
  } catch (const std::exception &e) {
    std::printf("\"error\": \"%s\"\n}\n", e.what());
    return 1;
  }
  std::printf("}\n");
  return 0;
}
