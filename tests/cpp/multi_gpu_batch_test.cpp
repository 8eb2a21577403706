// multi_gpu_batch_test.cpp -- the candidate loop of ModelTuner.getEnergyLandscape / AbstractCSA
// (CORE/optimization/ModelTuner.scala:131-153, AbstractCSA.scala:214-259) sharded over several GPUs from ONE
// process, through the C ABI only (include/dgp_b200.h), checked bit for bit against the single-context batch:
//   A. "one context per host thread per GPU": std::thread workers, each with its own dgp_ctx + dgp_data, looping
//      dgp_energy over their round-robin shard (the pattern a JVM thread pool would use);
//   B. dgp_energy_batch_multi over the same contexts from one call.
// With one visible GPU the contexts are two distinct contexts on that device, so the threading and sharding logic
// is exercised on the single-GPU test box as well.  Prints one JSON line; exit code 0 = all identical.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "dgp_b200.h"

static double lcg(uint64_t &s) {  // uniform in (-1, 1), deterministic
  s = s * 6364136223846793005ull + 1442695040888963407ull;
  return (double)((s >> 11) & ((1ull << 53) - 1)) / (double)(1ull << 52) - 1.0;
}

int main(int argc, char **argv) {
  const int64_t n = argc > 1 ? atoll(argv[1]) : 1500;
  const int d = 4, ncand = 13;
  std::vector<double> X((size_t)n * d), y((size_t)n);
  uint64_t seed = 2024;
  for (auto &v : X) v = 1.5 * lcg(seed);
  for (int64_t i = 0; i < n; ++i) y[i] = std::sin(X[i * d] + 0.5 * X[i * d + 1]) + 0.1 * lcg(seed);

  // every visible GPU; at least two contexts
  std::vector<dgp_ctx *> ctxs;
  int devices = 0;
  for (int dev = 0; dev < 16; ++dev) {
    dgp_ctx *c = nullptr;
    if (dgp_ctx_create(dev, &c) != DGP_OK) break;
    ctxs.push_back(c);
    ++devices;
  }
  if (ctxs.empty()) {
    std::printf("{\"error\": \"no usable GPU\"}\n");
    return 2;
  }
  if (ctxs.size() == 1) {
    dgp_ctx *c = nullptr;
    if (dgp_ctx_create(0, &c) != DGP_OK) return 2;
    ctxs.push_back(c);
  }
  const int nc = (int)ctxs.size();
  std::vector<dgp_data *> data(nc, nullptr);
  for (int c = 0; c < nc; ++c)
    if (dgp_data_create(ctxs[c], X.data(), n, d, DGP_ROW_MAJOR, y.data(), nullptr, &data[c]) != DGP_OK) {
      std::printf("{\"error\": \"data_create: %s\"}\n", dgp_last_error(ctxs[c]));
      return 2;
    }

  // candidates: SEKernel(bandwidth, amplitude) + DiracKernel(noise); one of them not positive definite (NaN bandwidth)
  std::vector<double> hyp((size_t)ncand * 2);
  std::vector<dgp_kernel_spec> specs(ncand);
  for (int i = 0; i < ncand; ++i) {
    hyp[2 * i] = (i == 5) ? NAN : 0.8 + 0.25 * i;
    hyp[2 * i + 1] = 1.0 + 0.05 * i;
    specs[i].kind = DGP_KERNEL_SE;
    specs[i].n_hyp = 2;
    specs[i].grad_mode = DGP_GRAD_REFERENCE;
    specs[i].reserved = 0;
    specs[i].hyp = &hyp[2 * i];
    specs[i].noise = 0.05 + 0.02 * i;
  }

  std::vector<double> e_ref(ncand), e_thr(ncand), e_multi(ncand);
  std::vector<int32_t> s_ref(ncand), s_thr(ncand), s_multi(ncand);
  if (dgp_energy_batch(ctxs[0], data[0], specs.data(), ncand, e_ref.data(), s_ref.data()) != DGP_OK) return 3;

  // A. one context per thread
  std::vector<std::thread> workers;
  for (int c = 0; c < nc; ++c)
    workers.emplace_back([&, c] {
      for (int i = c; i < ncand; i += nc) {
        s_thr[i] = dgp_energy(ctxs[c], data[c], &specs[i], &e_thr[i]);
      }
    });
  for (auto &w : workers) w.join();

  // B. one call
  const int32_t rc = dgp_energy_batch_multi(ctxs.data(), data.data(), nc, specs.data(), ncand, e_multi.data(), s_multi.data());

  int mismatches = rc == DGP_OK ? 0 : 1000;
  for (int i = 0; i < ncand; ++i) {
    const bool same_thr = std::memcmp(&e_ref[i], &e_thr[i], sizeof(double)) == 0 || (std::isinf(e_ref[i]) && std::isinf(e_thr[i]));
    const bool same_multi = std::memcmp(&e_ref[i], &e_multi[i], sizeof(double)) == 0;
    if (!same_thr || !same_multi || s_ref[i] != s_multi[i] || s_ref[i] != s_thr[i]) ++mismatches;
  }
  // argument checking of the multi entry point
  int32_t bad_args = 0;
  {
    dgp_ctx *twice[2] = {ctxs[0], ctxs[0]};
    const dgp_data *dd[2] = {data[0], data[0]};
    if (dgp_energy_batch_multi(twice, dd, 2, specs.data(), ncand, e_multi.data(), s_multi.data()) != DGP_ERR_ARG) ++bad_args;
    if (dgp_energy_batch_multi(nullptr, dd, 2, specs.data(), ncand, e_multi.data(), s_multi.data()) != DGP_ERR_ARG) ++bad_args;
  }
  std::printf("{\"devices\": %d, \"contexts\": %d, \"candidates\": %d, \"n\": %lld, \"mismatches\": %d, \"bad_arg_checks\": %d, "
              "\"non_pd_status\": %d, \"e0\": %.17g}\n",
              devices, nc, ncand, (long long)n, mismatches, bad_args, s_ref[5], e_ref[0]);
  for (int c = 0; c < nc; ++c) {
    dgp_data_destroy(ctxs[c], data[c]);
    dgp_ctx_destroy(ctxs[c]);
  }
  return (mismatches == 0 && bad_args == 0 && s_ref[5] == DGP_ERR_NOT_PD) ? 0 : 1;
}
