// dynaml_b200.hpp -- header-only C++17 host-side mirror of the dynaml-core classes on the GP hot path,
// layered on the C ABI of dgp_b200.h.  The reference's host code is Scala on the JVM, which cannot be
// built in this image; this is the compiled-language twin of the Scala shim described in
// same class names, hyper-parameter names, argument meaning and error behaviour
// (CORE = dynaml-core/src/main/scala/io/github/tailhq/dynaml):
//
//   LocalScalarKernel, SEKernel, RBFKernel, MahalanobisKernel, GenericMaternKernel, PeriodicKernel, DiracKernel
//       CORE/kernels/{CovarianceFunction,LocalScalarKernel,RBFKernel,GenericMaternKernel,PeriodicKernel,DiracKernel}.scala
//   GPRegression (energy, gradEnergy, predictiveDistribution, predictionWithErrorBars, persist)
//       CORE/models/gp/{AbstractGPRegressionModel,GPRegression}.scala
//   GridSearch, CoupledSimulatedAnnealing, GradBasedGlobalOptimizer
//       CORE/optimization/{ModelTuner,GridSearch,AbstractCSA,CoupledSimulatedAnnealing,GradBasedGlobalOptimizer}.scala
//
// Every number comes from libdgp_b200.so; there is no CPU fallback.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <limits>
#include <map>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "dgp_b200.h"

namespace dynaml {

using Config = std::map<std::string, double>;        // Map[String, Double]
using Options = std::map<std::string, std::string>;  // Map[String, String]
using Vec = std::vector<double>;                     // DenseVector[Double]

struct DenseMatrix {  // Breeze DenseMatrix[Double]: column-major, offset 0
  int64_t rows = 0, cols = 0;
  std::vector<double> data;
  double operator()(int64_t i, int64_t j) const { return data[(size_t)(i + j * rows)]; }
};

// PartitionedPSDMatrix (CORE/algebra/PartitionedMatrix.scala:248-259): the blocks on / below the block diagonal;
// the blocks above it are their transposes.  PartitionedMatrix: every block given.
struct PartitionedMatrix {
  int64_t rows = 0, cols = 0, rowBlocks = 0, colBlocks = 0;
  bool symmetric = false;  // true: only blocks (i, j), i >= j, are stored
  std::map<std::pair<int64_t, int64_t>, DenseMatrix> blocks;
  DenseMatrix toDenseMatrix() const {
    DenseMatrix out;
    out.rows = rows;
    out.cols = cols;
    out.data.assign((size_t)rows * cols, 0.0);
    std::vector<int64_t> r0(rowBlocks + 1, 0), c0(colBlocks + 1, 0);
    for (const auto &kv : blocks) {
      if (kv.first.second == 0) r0[kv.first.first + 1] = kv.second.rows;
      if (kv.first.first == kv.first.second || !symmetric) c0[kv.first.second + 1] = kv.second.cols;
    }
    for (int64_t i = 0; i < rowBlocks; ++i) r0[i + 1] += r0[i];
    for (int64_t j = 0; j < colBlocks; ++j) c0[j + 1] += c0[j];
    for (const auto &kv : blocks) {
      const DenseMatrix &b = kv.second;
      for (int64_t c = 0; c < b.cols; ++c)
        for (int64_t r = 0; r < b.rows; ++r) {
          out.data[(size_t)(r0[kv.first.first] + r + (c0[kv.first.second] + c) * rows)] = b(r, c);
          if (symmetric && kv.first.first != kv.first.second)
            out.data[(size_t)(c0[kv.first.second] + c + (r0[kv.first.first] + r) * rows)] = b(r, c);
        }
    }
    return out;
  }
};

class DgpError : public std::runtime_error {
 public:
  DgpError(int32_t st, const std::string &m) : std::runtime_error(m), status(st) {}
  int32_t status;
};

// One GPU (dgp_ctx); not thread-safe, like the reference model.
class Context {
 public:
  explicit Context(int device = 0) {
    if (dgp_ctx_create(device, &h_) != DGP_OK)
      throw DgpError(DGP_ERR_CUDA, "dgp_ctx_create failed: a B200 (sm_100) GPU is required; there is no CPU fallback");
  }
  ~Context() { dgp_ctx_destroy(h_); }
  Context(const Context &) = delete;
  Context &operator=(const Context &) = delete;
  dgp_ctx *get() const { return h_; }
  void check(int32_t st) const {
    if (st != DGP_OK) throw DgpError(st, dgp_last_error(h_));
  }
  static std::shared_ptr<Context> &global() {
    static std::shared_ptr<Context> c = std::make_shared<Context>(0);
    return c;
  }

 private:
  dgp_ctx *h_ = nullptr;
};

inline std::vector<double> pack(const std::vector<Vec> &pts) {
  const size_t d = pts.empty() ? 0 : pts[0].size();
  std::vector<double> out;
  out.reserve(pts.size() * d);
  for (const Vec &p : pts) {
    if (p.size() != d) throw std::invalid_argument("points disagree on dimension");
    out.insert(out.end(), p.begin(), p.end());
  }
  return out;
}

// ---- covariance functions ----------------------------------------------------------------------------
class LocalScalarKernel {
 public:
  virtual ~LocalScalarKernel() = default;
  std::vector<std::string> hyper_parameters;
  std::vector<std::string> blocked_hyper_parameters;
  Config state;
  int grad_mode = DGP_GRAD_REFERENCE;

  void block(const std::vector<std::string> &h) { blocked_hyper_parameters = h; }
  void block_all_hyper_parameters() { blocked_hyper_parameters = hyper_parameters; }
  std::vector<std::string> effective_hyper_parameters() const {
    std::vector<std::string> out;
    for (const auto &h : hyper_parameters)
      if (std::find(blocked_hyper_parameters.begin(), blocked_hyper_parameters.end(), h) == blocked_hyper_parameters.end())
        out.push_back(h);
    return out;
  }
  Config effective_state() const {
    Config out;
    for (const auto &h : effective_hyper_parameters()) out[h] = state.at(h);
    return out;
  }
  // CovarianceFunction.scala:37-44: all effective keys must be present
  virtual void setHyperParameters(const Config &h) {
    for (const auto &k : effective_hyper_parameters())
      if (!h.count(k)) throw std::invalid_argument("All hyper parameters must be contained in the arguments");
    for (const auto &k : effective_hyper_parameters()) state[k] = h.at(k);
  }

  virtual int device_kind() const = 0;
  virtual Vec device_hyp() const {
    Vec v;
    for (const auto &h : hyper_parameters) v.push_back(state.at(h));
    return v;
  }

  double evaluate(const Vec &x, const Vec &y) const { return buildCrossKernelMatrix({x}, {y})(0, 0); }

  // LocalScalarKernel.scala:39-44, 160-161 -> SVMKernel.buildPartitionedKernelMatrix (SVMKernel.scala:181-219): blocks of
  // rowBlocking x colBlocking, block row >= block column, cut from the points resident on the device.
  int64_t rowBlocking = 1000, colBlocking = 1000;
  void setBlockSizes(int64_t r, int64_t c) { rowBlocking = r; colBlocking = c; }
  PartitionedMatrix buildBlockedKernelMatrix(const std::vector<Vec> &data, double noise = 0.0, bool add_noise = false) const {
    if (rowBlocking != colBlocking) throw std::invalid_argument("buildBlockedKernelMatrix: equal block sizes are required");
    const auto ctx = Context::global();
    Vec hyp = device_hyp();
    dgp_kernel_spec sp{device_kind(), (int32_t)hyp.size(), grad_mode, 0, hyp.data(), noise};
    std::vector<double> X = pack(data), y(data.size(), 0.0);
    const int64_t n = (int64_t)data.size(), nb = (n + rowBlocking - 1) / rowBlocking;
    dgp_data *h = nullptr;
    ctx->check(dgp_data_create(ctx->get(), X.data(), n, (int32_t)data[0].size(), DGP_ROW_MAJOR, y.data(), nullptr, &h));
    PartitionedMatrix P;
    P.rows = P.cols = n;
    P.rowBlocks = P.colBlocks = nb;
    P.symmetric = true;
    for (int64_t i = 0; i < nb; ++i)
      for (int64_t j = 0; j <= i; ++j) {
        DenseMatrix b;
        b.rows = std::min(rowBlocking, n - i * rowBlocking);
        b.cols = std::min(colBlocking, n - j * colBlocking);
        b.data.resize((size_t)b.rows * b.cols);
        const int32_t st = dgp_data_kernel_block(ctx->get(), h, &sp, i * rowBlocking, b.rows, j * colBlocking, b.cols,
                                                 add_noise ? 1 : 0, b.data.data(), b.rows);
        if (st != DGP_OK) {
          dgp_data_destroy(ctx->get(), h);
          ctx->check(st);
        }
        P.blocks[{i, j}] = std::move(b);
      }
    dgp_data_destroy(ctx->get(), h);
    return P;
  }
  // SVMKernel.buildKernelGradMatrix (SVMKernel.scala:107-145): "kernel-matrix" and one matrix per hyper-parameter
  // (device order: this kernel's hyper_parameters, then "noiseLevel" when a noise level is given).
  std::map<std::string, DenseMatrix> buildKernelGradMatrix(const std::vector<Vec> &data, double noise = 0.0,
                                                           bool add_noise = false) const {
    const auto ctx = Context::global();
    Vec hyp = device_hyp();
    dgp_kernel_spec sp{device_kind(), (int32_t)hyp.size(), grad_mode, 0, hyp.data(), noise};
    const int ng = dgp_spec_grad_len(&sp);
    const int64_t n = (int64_t)data.size();
    std::vector<double> X = pack(data), out((size_t)(1 + ng) * n * n);
    ctx->check(dgp_kernel_grad_matrix(ctx->get(), &sp, X.data(), n, nullptr, 0, (int32_t)data[0].size(), DGP_ROW_MAJOR,
                                      add_noise ? 1 : 0, out.data()));
    std::vector<std::string> names = {"kernel-matrix"};
    for (const auto &h : grad_names()) names.push_back(h);
    names.push_back("noiseLevel");
    std::map<std::string, DenseMatrix> res;
    for (int q = 0; q <= ng && q < (int)names.size(); ++q) {
      DenseMatrix m;
      m.rows = m.cols = n;
      m.data.assign(out.begin() + (size_t)q * n * n, out.begin() + (size_t)(q + 1) * n * n);
      res[names[q]] = std::move(m);
    }
    return res;
  }
  virtual std::vector<std::string> grad_names() const { return hyper_parameters; }

  DenseMatrix buildKernelMatrix(const std::vector<Vec> &data, double noise = 0.0, bool add_noise = false) const {
    const auto ctx = Context::global();
    Vec hyp = device_hyp();
    dgp_kernel_spec sp{device_kind(), (int32_t)hyp.size(), grad_mode, 0, hyp.data(), noise};
    std::vector<double> X = pack(data);
    DenseMatrix K;
    K.rows = K.cols = (int64_t)data.size();
    K.data.resize((size_t)K.rows * K.cols);
    ctx->check(dgp_kernel_matrix(ctx->get(), &sp, X.data(), K.rows, nullptr, 0, (int32_t)data[0].size(), DGP_ROW_MAJOR,
                                 K.data.data(), K.rows, add_noise ? 1 : 0));
    return K;
  }
  DenseMatrix buildCrossKernelMatrix(const std::vector<Vec> &d1, const std::vector<Vec> &d2) const {
    const auto ctx = Context::global();
    Vec hyp = device_hyp();
    dgp_kernel_spec sp{device_kind(), (int32_t)hyp.size(), grad_mode, 0, hyp.data(), 0.0};
    std::vector<double> X1 = pack(d1), X2 = pack(d2);
    DenseMatrix K;
    K.rows = (int64_t)d1.size();
    K.cols = (int64_t)d2.size();
    K.data.resize((size_t)K.rows * K.cols);
    ctx->check(dgp_kernel_matrix(ctx->get(), &sp, X1.data(), K.rows, X2.data(), K.cols, (int32_t)d1[0].size(),
                                 DGP_ROW_MAJOR, K.data.data(), K.rows, 0));
    return K;
  }
};

class RBFKernel : public LocalScalarKernel {
 public:
  explicit RBFKernel(double bandwidth = 1.0) {
    hyper_parameters = {"bandwidth"};
    state = {{"bandwidth", bandwidth}};
  }
  int device_kind() const override { return DGP_KERNEL_RBF; }
};

class SEKernel : public LocalScalarKernel {
 public:
  explicit SEKernel(double band = 1.0, double h = 2.0) {
    hyper_parameters = {"bandwidth", "amplitude"};
    state = {{"bandwidth", band}, {"amplitude", h}};
  }
  int device_kind() const override { return DGP_KERNEL_SE; }
};

class MahalanobisKernel : public LocalScalarKernel {
 public:
  explicit MahalanobisKernel(const Vec &band, double h = 2.0, int mode = DGP_GRAD_REFERENCE) {
    hyper_parameters.push_back("MahalanobisAmplitude");
    state["MahalanobisAmplitude"] = h;
    for (size_t i = 0; i < band.size(); ++i) {
      hyper_parameters.push_back("MahalanobisBandwidth_" + std::to_string(i));
      state[hyper_parameters.back()] = band[i];
    }
    grad_mode = mode;
  }
  int device_kind() const override { return DGP_KERNEL_ARD_SE; }
};

class GenericMaternKernel : public LocalScalarKernel {
 public:
  explicit GenericMaternKernel(double l, int p = 2) {
    hyper_parameters = {"p", "l"};
    blocked_hyper_parameters = {"p"};
    state = {{"p", (double)p}, {"l", l}};
  }
  void setHyperParameters(const Config &h) override {  // GenericMaternKernel.scala:48-53
    LocalScalarKernel::setHyperParameters(h);
    if (h.count("p")) state["p"] = std::floor(std::fabs(h.at("p")));
  }
  int device_kind() const override { return DGP_KERNEL_MATERN; }
};

class PeriodicKernel : public LocalScalarKernel {
 public:
  explicit PeriodicKernel(double s = 1.0, double ls = 1.0, double freq = 1.0) {
    hyper_parameters = {"sigma", "lengthscale", "frequency"};
    state = {{"sigma", s}, {"lengthscale", ls}, {"frequency", freq}};
  }
  int device_kind() const override { return DGP_KERNEL_PERIODIC; }
};

// further stationary kernels of the same builder API
class CauchyKernel : public LocalScalarKernel {  // CORE/kernels/CauchyKernel.scala
 public:
  explicit CauchyKernel(double si = 1.0) {
    hyper_parameters = {"sigma"};
    state = {{"sigma", si}};
  }
  int device_kind() const override { return DGP_KERNEL_CAUCHY; }
};

class LaplacianKernel : public LocalScalarKernel {  // CORE/kernels/LaplacianKernel.scala
 public:
  explicit LaplacianKernel(double be = 1.0) {
    hyper_parameters = {"beta"};
    state = {{"beta", be}};
  }
  int device_kind() const override { return DGP_KERNEL_LAPLACIAN; }
};

class RationalQuadraticKernel : public LocalScalarKernel {  // CORE/kernels/RationalQuadraticKernel.scala
 public:
  explicit RationalQuadraticKernel(double shape = 1.0, double l = 1.0) {
    hyper_parameters = {"mu", "l"};
    state = {{"mu", shape}, {"l", l}};
  }
  int device_kind() const override { return DGP_KERNEL_RATQUAD; }
};

class TStudentKernel : public LocalScalarKernel {  // CORE/kernels/TStudentKernel.scala
 public:
  explicit TStudentKernel(double degree = 1.0) {
    hyper_parameters = {"d"};
    state = {{"d", degree}};
  }
  int device_kind() const override { return DGP_KERNEL_TSTUDENT; }
};

class PolynomialKernel : public LocalScalarKernel {  // CORE/kernels/PolynomialKernel.scala:9-45
 public:
  explicit PolynomialKernel(int degree = 2, double offset = 1.0) {
    hyper_parameters = {"degree", "offset"};
    state = {{"degree", (double)degree}, {"offset", offset}};
  }
  void setHyperParameters(const Config &h) override {
    LocalScalarKernel::setHyperParameters(h);
    if (h.count("offset")) state["offset"] = std::fabs(h.at("offset"));
  }
  int device_kind() const override { return DGP_KERNEL_POLYNOMIAL; }
};

class FBMKernel : public LocalScalarKernel {  // CORE/kernels/FBMKernel.scala:70-122
 public:
  explicit FBMKernel(double hurst = 0.75) {
    hyper_parameters = {"hurst"};
    state = {{"hurst", hurst}};
  }
  int device_kind() const override { return DGP_KERNEL_FBM; }
};

class DiracKernel : public LocalScalarKernel {
 public:
  explicit DiracKernel(double noiseLevel = 1.0) {
    hyper_parameters = {"noiseLevel"};
    state = {{"noiseLevel", noiseLevel}};
  }
  int device_kind() const override { return DGP_KERNEL_DIRAC; }
};

// ---- kernel algebra (LocalScalarKernel.scala:53-91) -------------------------------------------------------
// k1 + k2, k1 * k2 and k * c reduce to  sum_t coef_t prod_f leaf_f^e_tf ; device_hyp() is then the flat
// DGP_KERNEL_COMPOSITE program of dgp_b200.h.  Hyper-parameter names carry one "<Class@id>/" prefix per
// combinator and configurations are routed by prefix, as AdditiveCovariance.scala:34-58 does.
class CompositeCovariance : public LocalScalarKernel {
 public:
  using Term = std::pair<double, std::map<int, int>>;  // coefficient, leaf position -> exponent
  using KernelPtr = std::shared_ptr<LocalScalarKernel>;

  int device_kind() const override { return DGP_KERNEL_COMPOSITE; }
  Vec device_hyp() const override {
    std::vector<const LocalScalarKernel *> lv;
    leaves(lv);
    std::vector<Term> ts = terms();
    if (lv.empty() || lv.size() > DGP_MAX_FACTORS || ts.empty() || ts.size() > DGP_MAX_TERMS)
      throw std::invalid_argument("a composite kernel takes 1..32 base kernels and 1..64 terms (no CPU fallback)");
    Vec prog{(double)lv.size(), (double)ts.size()};
    for (const auto *k : lv) {
      Vec h = k->device_hyp();
      prog.push_back((double)k->device_kind());
      prog.push_back((double)k->grad_mode);
      prog.push_back((double)h.size());
      prog.insert(prog.end(), h.begin(), h.end());
    }
    for (const auto &t : ts) {
      prog.push_back(t.first);
      for (size_t f = 0; f < lv.size(); ++f) {
        auto it = t.second.find((int)f);
        prog.push_back(it == t.second.end() ? 0.0 : (double)it->second);
      }
    }
    return prog;
  }
  virtual void leaves(std::vector<const LocalScalarKernel *> &out) const = 0;
  virtual std::vector<Term> terms() const = 0;

  static void leaves_of(const LocalScalarKernel &k, std::vector<const LocalScalarKernel *> &out) {
    if (auto *c = dynamic_cast<const CompositeCovariance *>(&k)) c->leaves(out);
    else out.push_back(&k);
  }
  static std::vector<Term> terms_of(const LocalScalarKernel &k) {
    if (auto *c = dynamic_cast<const CompositeCovariance *>(&k)) return c->terms();
    return {Term{1.0, {{0, 1}}}};
  }
  static std::string id_of(const LocalScalarKernel &k, const char *cls) {
    char buf[96];
    snprintf(buf, sizeof buf, "%s@%llx", cls, (unsigned long long)(uintptr_t)&k & 0xffffffffull);
    return buf;
  }
};

class PairCovariance : public CompositeCovariance {
 public:
  PairCovariance(KernelPtr first, KernelPtr other, const char *fcls, const char *scls)
      : firstKernel(std::move(first)), otherKernel(std::move(other)) {
    fID = id_of(*firstKernel, fcls);
    sID = id_of(*otherKernel, scls);
    for (const auto &h : firstKernel->hyper_parameters) hyper_parameters.push_back(fID + "/" + h);
    for (const auto &h : otherKernel->hyper_parameters) hyper_parameters.push_back(sID + "/" + h);
    for (const auto &h : firstKernel->blocked_hyper_parameters) blocked_hyper_parameters.push_back(fID + "/" + h);
    for (const auto &h : otherKernel->blocked_hyper_parameters) blocked_hyper_parameters.push_back(sID + "/" + h);
    for (const auto &kv : firstKernel->state) state[fID + "/" + kv.first] = kv.second;
    for (const auto &kv : otherKernel->state) state[sID + "/" + kv.first] = kv.second;
  }
  KernelPtr firstKernel, otherKernel;
  std::string fID, sID;

  void setHyperParameters(const Config &h) override {
    Config c1, c2;
    for (const auto &kv : h) {
      if (kv.first.rfind(fID + "/", 0) == 0) c1[kv.first.substr(fID.size() + 1)] = kv.second;
      if (kv.first.rfind(sID + "/", 0) == 0) c2[kv.first.substr(sID.size() + 1)] = kv.second;
    }
    firstKernel->setHyperParameters(c1);
    otherKernel->setHyperParameters(c2);
    LocalScalarKernel::setHyperParameters(h);
  }
  void leaves(std::vector<const LocalScalarKernel *> &out) const override {
    leaves_of(*firstKernel, out);
    leaves_of(*otherKernel, out);
  }

 protected:
  std::vector<Term> shifted_other() const {
    std::vector<const LocalScalarKernel *> lv;
    leaves_of(*firstKernel, lv);
    std::vector<Term> out;
    for (const auto &t : terms_of(*otherKernel)) {
      Term s{t.first, {}};
      for (const auto &fe : t.second) s.second[fe.first + (int)lv.size()] = fe.second;
      out.push_back(s);
    }
    return out;
  }
};

class AdditiveCovariance : public PairCovariance {  // AdditiveCovariance.scala:27-93
 public:
  AdditiveCovariance(KernelPtr first, KernelPtr other)
      : PairCovariance(std::move(first), std::move(other), "Kernel", "Kernel") {}
  std::vector<Term> terms() const override {
    std::vector<Term> out = terms_of(*firstKernel), o = shifted_other();
    out.insert(out.end(), o.begin(), o.end());
    return out;
  }
};

class MultiplicativeCovariance : public PairCovariance {  // MultiplicativeCovariance.scala:11-90
 public:
  MultiplicativeCovariance(KernelPtr first, KernelPtr other)
      : PairCovariance(std::move(first), std::move(other), "Kernel", "Kernel") {}
  std::vector<Term> terms() const override {
    std::vector<Term> out;
    for (const auto &a : terms_of(*firstKernel))
      for (const auto &b : shifted_other()) {
        Term t{a.first * b.first, a.second};
        for (const auto &fe : b.second) t.second[fe.first] += fe.second;
        out.push_back(t);
      }
    return out;
  }
};

class ScaledKernel : public CompositeCovariance {  // kernel * c, LocalScalarKernel.scala:68-91
 public:
  ScaledKernel(KernelPtr base_, double c_) : base(std::move(base_)), c(c_) {
    if (!(c > 0)) throw std::invalid_argument("Multiplicative constant applied on a kernel must be positive!");
    hyper_parameters = base->hyper_parameters;
    blocked_hyper_parameters = base->blocked_hyper_parameters;
    state = base->state;
  }
  KernelPtr base;
  double c;
  void setHyperParameters(const Config &h) override {
    base->setHyperParameters(h);
    LocalScalarKernel::setHyperParameters(h);
  }
  void leaves(std::vector<const LocalScalarKernel *> &out) const override { leaves_of(*base, out); }
  std::vector<Term> terms() const override {
    std::vector<Term> out = terms_of(*base);
    for (auto &t : out) t.first *= c;
    return out;
  }
};

inline std::shared_ptr<LocalScalarKernel> operator+(std::shared_ptr<LocalScalarKernel> a,
                                                    std::shared_ptr<LocalScalarKernel> b) {
  return std::make_shared<AdditiveCovariance>(std::move(a), std::move(b));
}
inline std::shared_ptr<LocalScalarKernel> operator*(std::shared_ptr<LocalScalarKernel> a,
                                                    std::shared_ptr<LocalScalarKernel> b) {
  return std::make_shared<MultiplicativeCovariance>(std::move(a), std::move(b));
}
inline std::shared_ptr<LocalScalarKernel> operator*(std::shared_ptr<LocalScalarKernel> a, double c) {
  return std::make_shared<ScaledKernel>(std::move(a), c);
}

// ---- results -------------------------------------------------------------------------------------------
struct MultGaussianPRV {
  Vec mu;
  DenseMatrix covariance;
  std::function<std::pair<Vec, Vec>(double)> confidenceInterval;  // BlockedMultiVariateGaussian.scala:58-68
};

// ---- model -----------------------------------------------------------------------------------------------
class GPRegression {
 public:
  using Data = std::vector<std::pair<Vec, double>>;
  using Mean = std::function<double(const Vec &)>;

  GPRegression(std::shared_ptr<LocalScalarKernel> cov, std::shared_ptr<DiracKernel> noise, Data trainingdata,
               Mean meanFunc = [](const Vec &) { return 0.0; })
      : covariance(std::move(cov)), noiseModel(std::move(noise)), g(std::move(trainingdata)), mean(std::move(meanFunc)),
        ctx_(Context::global()) {
    if (g.empty()) throw std::invalid_argument("training data is empty");
    covariance->device_kind();
    hyper_parameters = covariance->hyper_parameters;
    hyper_parameters.insert(hyper_parameters.end(), noiseModel->hyper_parameters.begin(),
                            noiseModel->hyper_parameters.end());
    refresh_state();
  }
  virtual ~GPRegression() {
    if (model_) dgp_model_destroy(ctx_->get(), model_);
    if (data_) dgp_data_destroy(ctx_->get(), data_);
  }
  GPRegression(const GPRegression &) = delete;

  std::shared_ptr<LocalScalarKernel> covariance;
  std::shared_ptr<DiracKernel> noiseModel;
  std::vector<std::string> hyper_parameters;
  const Config &_current_state() const { return current_state; }

  void setState(const Config &s) {  // AbstractGPRegressionModel.scala:117-128
    Config c, n;
    for (const auto &kv : s) {
      if (has(covariance->hyper_parameters, kv.first)) c[kv.first] = kv.second;
      if (has(noiseModel->hyper_parameters, kv.first)) n[kv.first] = kv.second;
    }
    covariance->setHyperParameters(c);
    noiseModel->setHyperParameters(n);
    refresh_state();
  }

  // 0.5*(r'K^-1 r + sum log L_ii + n log 2pi); +Inf when K is not PD (:148-189, 516-535)
  double energy(const Config &h, const Options & = {}) {
    setState(h);
    Vec hyp = covariance->device_hyp();
    dgp_kernel_spec sp = spec(hyp);
    double e = 0.0;
    const int32_t st = dgp_energy(ctx_->get(), device_data(), &sp, &e);
    if (st == DGP_ERR_NOT_PD) return std::numeric_limits<double>::infinity();
    ctx_->check(st);
    return e;
  }

  std::vector<double> energyBatch(const std::vector<Config> &hs) {
    std::vector<Vec> hyps;
    std::vector<dgp_kernel_spec> specs;
    for (const Config &h : hs) {
      setState(h);
      hyps.push_back(covariance->device_hyp());
    }
    for (size_t i = 0; i < hs.size(); ++i) {
      setState(hs[i]);
      specs.push_back(spec(hyps[i]));
    }
    std::vector<double> e(hs.size());
    std::vector<int32_t> status(hs.size());
    if (!hs.empty())
      ctx_->check(dgp_energy_batch(ctx_->get(), device_data(), specs.data(), (int32_t)specs.size(), e.data(), status.data()));
    return e;
  }

  // tr((alpha alpha' - K^-1) dK/dtheta) per effective hyper-parameter; NaN map on failure (:196-267)
  Config gradEnergy(const Config &h) {
    covariance->setHyperParameters(h);
    noiseModel->setHyperParameters(h);
    refresh_state();
    Vec hyp = covariance->device_hyp();
    dgp_kernel_spec sp = spec(hyp);
    const int32_t ng = dgp_spec_grad_len(&sp);
    if (ng < 1) throw std::invalid_argument("malformed kernel spec");
    std::vector<double> gr((size_t)ng);
    double e = 0.0;
    const int32_t st = dgp_energy_grad(ctx_->get(), device_data(), &sp, &e, gr.data());
    if (st != DGP_ERR_NOT_PD) ctx_->check(st);
    Config out;
    const auto &names = covariance->hyper_parameters;
    for (const auto &k : covariance->effective_hyper_parameters())
      out[k] = gr[(size_t)(std::find(names.begin(), names.end(), k) - names.begin())];
    for (const auto &k : noiseModel->effective_hyper_parameters()) out[k] = gr.back();
    return out;
  }

  void persist(const Config &state) {  // :399-413; keeps L and alpha resident instead of K
    setState(state);
    drop_model();
    Vec hyp = covariance->device_hyp();
    dgp_kernel_spec sp = spec(hyp);
    ctx_->check(dgp_fit(ctx_->get(), device_data(), &sp, &model_));
    caching_ = true;
  }
  void unpersist() {
    drop_model();
    caching_ = false;
  }

  MultGaussianPRV predictiveDistribution(const std::vector<Vec> &test) {  // :304-357
    MultGaussianPRV out;
    run_predict(test, 1.0, &out.mu, &out.covariance, nullptr, nullptr);
    out.confidenceInterval = [this, test](double s) {
      Vec mu, lo, hi;
      run_predict(test, s, &mu, nullptr, &lo, &hi);
      return std::make_pair(lo, hi);
    };
    return out;
  }

  std::vector<std::tuple<Vec, double, double, double>> predictionWithErrorBars(const std::vector<Vec> &test, int sigma) {
    Vec mu, lo, hi;
    run_predict(test, (double)sigma, &mu, nullptr, &lo, &hi);
    std::vector<std::tuple<Vec, double, double, double>> out;
    for (size_t i = 0; i < test.size(); ++i) out.emplace_back(test[i], mu[i], lo[i], hi[i]);
    return out;
  }
  double predict(const Vec &point) { return std::get<1>(predictionWithErrorBars({point}, 1)[0]); }

 protected:
  // hooks of GPBasisFuncRegression (explicit basis functions); no-ops for the plain model
  virtual void after_data_create(dgp_data *, const std::vector<Vec> &) {}
  virtual void before_predict(dgp_model *, const std::vector<Vec> &) {}
  static bool has(const std::vector<std::string> &v, const std::string &k) {
    return std::find(v.begin(), v.end(), k) != v.end();
  }
  void refresh_state() {
    current_state = covariance->state;
    for (const auto &kv : noiseModel->state) current_state[kv.first] = kv.second;
  }
  dgp_kernel_spec spec(Vec &hyp) const {
    return dgp_kernel_spec{covariance->device_kind(), (int32_t)hyp.size(), covariance->grad_mode, 0, hyp.data(),
                           noiseModel->state.at("noiseLevel")};
  }
  dgp_data *device_data() {
    if (!data_) {
      std::vector<Vec> pts;
      std::vector<double> y, m;
      for (const auto &p : g) {
        pts.push_back(p.first);
        y.push_back(p.second);
        m.push_back(mean(p.first));
      }
      std::vector<double> X = pack(pts);
      ctx_->check(dgp_data_create(ctx_->get(), X.data(), (int64_t)g.size(), (int32_t)pts[0].size(), DGP_ROW_MAJOR, y.data(),
                                  m.data(), &data_));
      after_data_create(data_, pts);
    }
    return data_;
  }
  void drop_model() {
    if (model_) dgp_model_destroy(ctx_->get(), model_);
    model_ = nullptr;
  }
  void run_predict(const std::vector<Vec> &test, double sigma, Vec *mu, DenseMatrix *cov, Vec *lo, Vec *hi) {
    dgp_model *m = model_;
    const bool temp = !(model_ && caching_);
    if (temp) {
      Vec hyp = covariance->device_hyp();
      dgp_kernel_spec sp = spec(hyp);
      ctx_->check(dgp_fit(ctx_->get(), device_data(), &sp, &m));
    }
    const int64_t M = (int64_t)test.size();
    std::vector<double> Xs = pack(test), ms;
    for (const Vec &t : test) ms.push_back(mean(t));
    mu->assign((size_t)M, 0.0);
    if (cov) {
      cov->rows = cov->cols = M;
      cov->data.assign((size_t)M * M, 0.0);
    }
    if (lo) {
      lo->assign((size_t)M, 0.0);
      hi->assign((size_t)M, 0.0);
    }
    before_predict(m, test);
    const int32_t st = dgp_predict(ctx_->get(), m, Xs.data(), M, DGP_ROW_MAJOR, ms.data(), mu->data(),
                                   cov ? cov->data.data() : nullptr, M, lo ? lo->data() : nullptr,
                                   hi ? hi->data() : nullptr, sigma);
    if (temp) dgp_model_destroy(ctx_->get(), m);
    ctx_->check(st);
  }

  Data g;
  Mean mean;
  Config current_state;
  std::shared_ptr<Context> ctx_;
  dgp_data *data_ = nullptr;
  dgp_model *model_ = nullptr;
  bool caching_ = false;
};

// GPBasisFuncRegressionModel (CORE/models/gp/GPBasisFuncRegressionModel.scala:70-109): explicit basis functions for
// the trend with a Gaussian prior N(b, covB) on their coefficients.  mean(x) = basisFunc(x).b and every kernel matrix
// gains (lowB basisFunc(x)).(lowB basisFunc(y)), lowB = cholesky(covB) -- applied on the device as a rank-m update.
// gradEnergy is the inherited one of the reference, which leaves the feature-map term out (:196-253).
class GPBasisFuncRegression : public GPRegression {
 public:
  using Basis = std::function<Vec(const Vec &)>;
  GPBasisFuncRegression(std::shared_ptr<LocalScalarKernel> cov, std::shared_ptr<DiracKernel> noise, Data trainingdata,
                        Basis basisFunc, Vec b, const DenseMatrix &covB)
      : GPRegression(std::move(cov), std::move(noise), std::move(trainingdata), make_mean(basisFunc, b)),
        basis_(std::move(basisFunc)), m_((int)b.size()) {
    if (covB.rows != m_ || covB.cols != m_) throw std::invalid_argument("basis_param_prior: covariance must be m x m");
    low_.assign((size_t)m_ * m_, 0.0);  // lower Cholesky factor, row-major
    for (int j = 0; j < m_; ++j)
      for (int i = j; i < m_; ++i) {
        double s = covB(i, j);
        for (int k = 0; k < j; ++k) s -= low_[(size_t)i * m_ + k] * low_[(size_t)j * m_ + k];
        if (i == j) {
          if (!(s > 0.0)) throw std::invalid_argument("basis_param_prior: covariance is not positive definite");
          low_[(size_t)i * m_ + j] = std::sqrt(s);
        } else {
          low_[(size_t)i * m_ + j] = s / low_[(size_t)j * m_ + j];
        }
      }
  }

 protected:
  void after_data_create(dgp_data *d, const std::vector<Vec> &pts) override {
    std::vector<double> psi = features(pts);
    ctx_->check(dgp_data_set_basis(ctx_->get(), d, psi.data(), m_));
  }
  void before_predict(dgp_model *m, const std::vector<Vec> &test) override {
    std::vector<double> psi = features(test);
    ctx_->check(dgp_model_set_test_basis(ctx_->get(), m, psi.data(), (int64_t)test.size(), m_));
  }

 private:
  static Mean make_mean(Basis f, Vec b) {
    return [f, b](const Vec &x) {
      Vec phi = f(x);
      double s = 0.0;
      for (size_t i = 0; i < b.size(); ++i) s += phi[i] * b[i];
      return s;
    };
  }
  std::vector<double> features(const std::vector<Vec> &pts) const {  // row i = (lowB basisFunc(x_i))'
    std::vector<double> psi(pts.size() * (size_t)m_);
    for (size_t p = 0; p < pts.size(); ++p) {
      Vec phi = basis_(pts[p]);
      if ((int)phi.size() != m_) throw std::invalid_argument("basisFunc must return one feature per trend coefficient");
      for (int i = 0; i < m_; ++i) {
        double s = 0.0;
        for (int k = 0; k <= i; ++k) s += low_[(size_t)i * m_ + k] * phi[(size_t)k];
        psi[p * (size_t)m_ + i] = s;
      }
    }
    return psi;
  }
  Basis basis_;
  int m_;
  std::vector<double> low_;
};

// ---- tuners ------------------------------------------------------------------------------------------------
class ModelTuner {
 public:
  explicit ModelTuner(GPRegression &model) : system(model) {}
  GPRegression &system;
  ModelTuner &setGridSize(int s) { gridsize = s; return *this; }
  ModelTuner &setStepSize(double s) { step = s; return *this; }
  ModelTuner &setLogScale(bool t) { logarithmicScale = t; return *this; }

  // ModelTuner.scala:79-103 (key order: std::map iterates keys sorted; Scala's small Maps keep insertion order)
  std::vector<Config> getGrid(const Config &initialConfig) const {
    std::vector<std::string> keys;
    std::vector<Vec> vecs;
    for (const auto &kv : initialConfig) {
      keys.push_back(kv.first);
      Vec v;
      for (int i = 0; i < gridsize; ++i)
        v.push_back(logarithmicScale ? kv.second / std::exp((i + 1) * step) : kv.second - (i + 1) * step);
      vecs.push_back(v);
    }
    std::vector<Config> grid(1);
    for (size_t k = 0; k < keys.size(); ++k) {
      std::vector<Config> next;
      for (const Config &c : grid)
        for (double v : vecs[k]) {
          Config n = c;
          n[keys[k]] = v;
          next.push_back(n);
        }
      grid.swap(next);
    }
    return grid;
  }
  std::vector<std::pair<double, Config>> getEnergyLandscape(const Config &initialConfig) {
    std::vector<Config> grid = getGrid(initialConfig);
    std::vector<double> e = system.energyBatch(grid);  // the reference's serial grid.map(system.energy)
    std::vector<std::pair<double, Config>> out;
    for (size_t i = 0; i < grid.size(); ++i) out.emplace_back(e[i], grid[i]);
    return out;
  }

 protected:
  double step = 0.3;
  int gridsize = 3;
  bool logarithmicScale = false;
};

class GridSearch : public ModelTuner {
 public:
  using ModelTuner::ModelTuner;
  std::pair<GPRegression *, Config> optimize(const Config &initialConfig, const Options &options = {}) {
    std::map<double, Config> landscape;  // `.toMap` then `keys.min` (GridSearch.scala:47-49)
    for (auto &kv : getEnergyLandscape(initialConfig)) landscape[kv.first] = kv.second;
    Config best = landscape.begin()->second;
    auto it = options.find("persist");
    if (it != options.end() && (it->second == "true" || it->second == "1")) system.persist(best);
    return {&system, best};
  }
};

class CoupledSimulatedAnnealing : public ModelTuner {
 public:
  explicit CoupledSimulatedAnnealing(GPRegression &model, uint64_t seed = 0) : ModelTuner(model), rng(seed) {}
  CoupledSimulatedAnnealing &setMaxIterations(int m) { MAX_ITERATIONS = m; return *this; }
  CoupledSimulatedAnnealing &setVariant(const std::string &v) { variant = v; return *this; }
  double iTemp = 1.0, alpha = 0.05;

  static double couplingFactor(const std::string &v, const std::vector<double> &land, double Tacc) {  // AbstractCSA.scala:285-297
    double s = 0.0;
    if (v == "CSA-MuSA" || v == "CSA-BA") { for (double e : land) s += std::exp(-e / Tacc); return s; }
    if (v == "CSA-M" || v == "CSA-MwVC") { for (double e : land) s += std::exp(e / Tacc); return s; }
    return 1.0;
  }
  static double acceptanceProbability(const std::string &v, double e, double old, double gamma, double t) {  // :299-316
    if (v == "CSA-MuSA") return std::exp(-e / t) / (std::exp(-e / t) + gamma);
    if (v == "CSA-BA") return 1.0 - std::exp(-old / t) / gamma;
    if (v == "CSA-M" || v == "CSA-MwVC") return std::exp(old / t) / gamma;
    return gamma / (1.0 + std::exp((e - old) / t));
  }

  std::pair<GPRegression *, Config> optimize(const Config &initialConfig, const Options &options = {}) {
    auto landscape = getEnergyLandscape(initialConfig);
    const double m = std::pow((double)gridsize, (double)initialConfig.size());
    const double sigmaD = (variant == "CSA-MuSA" || variant == "CSA-BA") ? 0.99 : 0.99 * (m - 1) / (m * m);
    std::vector<double> probs;
    {
      std::vector<double> es;
      for (auto &kv : landscape) es.push_back(kv.first);
      const double g0 = couplingFactor(variant, es, iTemp);
      for (double e : es) probs.push_back(acceptanceProbability(variant, e, e, g0, iTemp));
    }
    double accPrev = iTemp;
    std::cauchy_distribution<double> cauchy(0.0, 1.0);
    std::uniform_real_distribution<double> unif(0.0, 1.0);
    for (int it = MAX_ITERATIONS; it >= 1; --it) {
      const double mutTemp = iTemp / it;
      double accTemp = iTemp / std::log(it + 1.0);
      if (variant == "CSA-MwVC") {
        double mean = 0.0, var = 0.0;
        for (double p : probs) mean += p / probs.size();
        for (double p : probs) var += (p - mean) * (p - mean) / (probs.size() - 1);
        accTemp = var < sigmaD ? accPrev * (1 - alpha) : accPrev * (1 + alpha);
      }
      double maxE = -std::numeric_limits<double>::infinity();
      for (auto &kv : landscape) maxE = std::max(maxE, kv.first);
      std::vector<double> shifted;
      for (auto &kv : landscape) shifted.push_back(kv.first - maxE);
      const double gamma = couplingFactor(variant, shifted, accTemp);
      std::vector<Config> proposals;
      for (auto &kv : landscape) {
        Config c;
        for (auto &p : kv.second) c[p.first] = std::fabs(p.second + mutTemp * cauchy(rng));  // :112-130
        proposals.push_back(c);
      }
      std::vector<double> e = system.energyBatch(proposals);  // one sweep = one device batch
      for (size_t i = 0; i < landscape.size(); ++i) {
        const double p = acceptanceProbability(variant, e[i] - maxE, landscape[i].first - maxE, gamma, accTemp);
        if (e[i] < landscape[i].first || unif(rng) <= p) landscape[i] = {e[i], proposals[i]};
        probs[i] = p;
      }
      accPrev = accTemp;
    }
    std::map<double, Config> as_map;
    for (auto &kv : landscape) as_map[kv.first] = kv.second;
    Config best = as_map.begin()->second;
    auto it = options.find("persist");
    if (it != options.end() && (it->second == "true" || it->second == "1")) system.persist(best);
    return {&system, best};
  }

 private:
  int MAX_ITERATIONS = 10;
  std::string variant = "CSA-MuSA";
  std::mt19937_64 rng;  // the reference's RNGs are unseeded globals; injected here
};

class GradBasedGlobalOptimizer {
 public:
  explicit GradBasedGlobalOptimizer(GPRegression &model) : system(model) {}
  GPRegression &system;
  // h <- |h - step * g| until ||g|| < tolerance or maxIterations (GradBasedGlobalOptimizer.scala:37-103)
  std::pair<GPRegression *, Config> optimize(const Config &initialConfig, double tolerance = 1e-4, double step = 0.005,
                                             int maxIterations = 50) {
    Config working = initialConfig;
    int count = 1;
    double gradNorm = 1.0;
    do {
      Config g = system.gradEnergy(working);
      gradNorm = 0.0;
      for (auto &kv : g) gradNorm += kv.second * kv.second;
      gradNorm = std::sqrt(gradNorm);
      Config next;
      auto wi = working.begin();
      auto gi = g.begin();
      for (; wi != working.end() && gi != g.end(); ++wi, ++gi) {  // positional zip of the two maps (:69)
        double gr = gi->second;
        if (gr == std::numeric_limits<double>::infinity()) gr = 1.0;
        else if (gr == -std::numeric_limits<double>::infinity()) gr = -1.0;
        next[wi->first] = std::fabs(wi->second - step * gr);
      }
      working = next;
      ++count;
    } while (count < maxIterations && gradNorm >= tolerance);
    return {&system, working};
  }
};

}  // namespace dynaml
