// TEST INFRASTRUCTURE — part of the CPU oracle, never linked into the product library.
//
// "mini-Ceres": a from-scratch restatement of the slice of the ceres-solver API
// that the reference calls, so that the reference's OWN sources
// (/root/reference/src/*.cc, *.hh) compile unmodified into oracle/_ref/ and so
// that the oracle restatement in oracle/ba_oracle.cc shares one CPU solver.
//
// Ceres itself is an un-vendored, un-versioned dependency of the reference
// (CMakeLists.txt:7 `find_package(Ceres REQUIRED)`) and is not installable in
// this image, so everything below is restated from the published behaviour of
// upstream ceres-solver 2.x [Ceres-upstream]; PARITY IS UNPINNED for the solver
// semantics (the reference ships no tests / golden vectors).
//
// API surface provided == API surface used by the reference:
//   ceres::DynamicAutoDiffCostFunction<F,Stride>  snavely_reprojection_error.hh:11-14,128-139
//   ceres::AutoDiffCostFunction<F,1,3,1>          hemisphere_radius.hh:33
//   ceres::Problem::AddResidualBlock (vector and variadic forms),
//   ceres::Problem::SetParameterBlockConstant     sfm.cc:48,51-62,92
//   ceres::Solver::Options / Summary, ceres::Solve, Summary::FullReport
//                                                 sfm.cc:66-74,94-102
//   google::InitGoogleLogging                     sfm.cc:79
//   Eigen::Matrix3d/Vector3d/Map (through this header, as upstream's does)
#ifndef ORACLE_CERES_SHIM_CERES_H_
#define ORACLE_CERES_SHIM_CERES_H_

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <iostream>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <utility>
#include <limits>
#include <vector>

#include "Eigen/Core"
#include "ceres/jet.h"

namespace google {
inline void InitGoogleLogging(const char*) {}
}  // namespace google

namespace ceres {

// ---------------------------------------------------------------- cost functions
class CostFunction {
 public:
  CostFunction() : num_residuals_(0) {}
  virtual ~CostFunction() {}
  // jacobians[i] (if non-null) is row-major num_residuals x parameter_block_sizes()[i].
  virtual bool Evaluate(double const* const* parameters, double* residuals,
                        double** jacobians) const = 0;
  const std::vector<int32_t>& parameter_block_sizes() const { return parameter_block_sizes_; }
  int num_residuals() const { return num_residuals_; }

 protected:
  std::vector<int32_t>* mutable_parameter_block_sizes() { return &parameter_block_sizes_; }
  void set_num_residuals(int n) { num_residuals_ = n; }

 private:
  std::vector<int32_t> parameter_block_sizes_;
  int num_residuals_;
};

// [Ceres-upstream loss_function.h]  rho(s) of the squared residual norm s: out = {rho, rho', rho''}.
// The reference passes NULL (sfm.cc:48) and carries `new ceres::CauchyLoss(0.5)` in a comment (:49).
class LossFunction {
 public:
  virtual ~LossFunction() {}
  virtual void Evaluate(double sq_norm, double out[3]) const = 0;
};

// [Ceres-upstream loss_function.cc]  rho(s) = b log(1 + s / b), b = a^2
class CauchyLoss : public LossFunction {
 public:
  explicit CauchyLoss(double a) : b_(a * a), c_(1.0 / b_) {}
  void Evaluate(double s, double rho[3]) const override {
    const double sum = 1.0 + s * c_;
    const double inv = 1.0 / sum;
    rho[0] = b_ * std::log(sum);
    rho[1] = std::max(std::numeric_limits<double>::min(), inv);
    rho[2] = -c_ * (inv * inv);
  }

 private:
  const double b_, c_;
};

// Evaluates the functor with Jet<double,Stride> in ceil(active/Stride) passes
// [Ceres-upstream dynamic_autodiff_cost_function.h]; the functor signature is
//   template <class T> bool operator()(T const* const* params, T* residuals) const.
template <typename CostFunctor, int Stride = 4>
class DynamicAutoDiffCostFunction : public CostFunction {
 public:
  explicit DynamicAutoDiffCostFunction(CostFunctor* functor) : functor_(functor) {}
  virtual ~DynamicAutoDiffCostFunction() {}

  void AddParameterBlock(int size) { mutable_parameter_block_sizes()->push_back(size); }
  void SetNumResiduals(int num_residuals) { set_num_residuals(num_residuals); }

  bool Evaluate(double const* const* parameters, double* residuals,
                double** jacobians) const override {
    if (jacobians == NULL) {
      return (*functor_)(parameters, residuals);
    }
    typedef Jet<double, Stride> JetT;
    const std::vector<int32_t>& sizes = parameter_block_sizes();
    const int num_blocks = static_cast<int>(sizes.size());
    int num_parameters = 0;
    for (int i = 0; i < num_blocks; ++i) num_parameters += sizes[i];

    std::vector<JetT> input_jets(num_parameters > 0 ? num_parameters : 1);
    std::vector<JetT> output_jets(num_residuals());
    std::vector<JetT*> jet_parameters(num_blocks, static_cast<JetT*>(NULL));

    // active (differentiated) scalars, in raw-parameter order
    std::vector<int> active_block, active_col;
    std::vector<int> raw_index;
    int cursor = 0;
    for (int i = 0; i < num_blocks; ++i) {
      jet_parameters[i] = &input_jets[0] + cursor;
      for (int j = 0; j < sizes[i]; ++j, ++cursor) {
        input_jets[cursor].a = parameters[i][j];
        if (jacobians[i] != NULL) {
          active_block.push_back(i);
          active_col.push_back(j);
          raw_index.push_back(cursor);
        }
      }
    }
    const int num_active = static_cast<int>(raw_index.size());
    const int num_strides = (num_active + Stride - 1) / Stride;
    if (num_strides == 0) {
      return (*functor_)(parameters, residuals);
    }
    for (int pass = 0; pass < num_strides; ++pass) {
      const int begin = pass * Stride;
      const int end = std::min(num_active, begin + Stride);
      for (int p = 0; p < num_parameters; ++p)
        for (int s = 0; s < Stride; ++s) input_jets[p].v[s] = 0.0;
      for (int a = begin; a < end; ++a) input_jets[raw_index[a]].v[a - begin] = 1.0;
      if (!(*functor_)(&jet_parameters[0], &output_jets[0])) return false;
      for (int a = begin; a < end; ++a) {
        const int blk = active_block[a];
        const int col = active_col[a];
        for (int k = 0; k < num_residuals(); ++k)
          jacobians[blk][k * sizes[blk] + col] = output_jets[k].v[a - begin];
      }
      if (pass == num_strides - 1)
        for (int k = 0; k < num_residuals(); ++k) residuals[k] = output_jets[k].a;
    }
    return true;
  }

 private:
  std::unique_ptr<CostFunctor> functor_;
};

namespace internal {
template <int... Ns>
struct StaticSum;
template <>
struct StaticSum<> {
  static const int value = 0;
};
template <int N, int... Ns>
struct StaticSum<N, Ns...> {
  static const int value = N + StaticSum<Ns...>::value;
};
template <typename F, typename T, size_t... I>
inline bool CallVariadic(const F& f, T const* const* p, T* r, std::index_sequence<I...>) {
  return f(p[I]..., r);
}
}  // namespace internal

// Sized autodiff cost function; functor signature
//   template <class T> bool operator()(const T* x0, ..., T* residuals) const.
template <typename CostFunctor, int kNumResiduals, int... Ns>
class AutoDiffCostFunction : public CostFunction {
 public:
  explicit AutoDiffCostFunction(CostFunctor* functor) : functor_(functor) {
    set_num_residuals(kNumResiduals);
    const int sizes[] = {Ns...};
    for (size_t i = 0; i < sizeof...(Ns); ++i) mutable_parameter_block_sizes()->push_back(sizes[i]);
  }
  virtual ~AutoDiffCostFunction() {}

  bool Evaluate(double const* const* parameters, double* residuals,
                double** jacobians) const override {
    const int kBlocks = sizeof...(Ns);
    if (jacobians == NULL) {
      return internal::CallVariadic(*functor_, parameters, residuals,
                                    std::make_index_sequence<sizeof...(Ns)>());
    }
    const int kTotal = internal::StaticSum<Ns...>::value;
    typedef Jet<double, kTotal> JetT;
    const int sizes[] = {Ns...};
    JetT x[kTotal];
    JetT y[kNumResiduals];
    const JetT* ptrs[sizeof...(Ns)];
    int cursor = 0;
    for (int i = 0; i < kBlocks; ++i) {
      ptrs[i] = x + cursor;
      for (int j = 0; j < sizes[i]; ++j, ++cursor) x[cursor] = JetT(parameters[i][j], cursor);
    }
    if (!internal::CallVariadic(*functor_, ptrs, y, std::make_index_sequence<sizeof...(Ns)>()))
      return false;
    for (int k = 0; k < kNumResiduals; ++k) residuals[k] = y[k].a;
    cursor = 0;
    for (int i = 0; i < kBlocks; ++i) {
      if (jacobians[i] != NULL)
        for (int k = 0; k < kNumResiduals; ++k)
          for (int j = 0; j < sizes[i]; ++j) jacobians[i][k * sizes[i] + j] = y[k].v[cursor + j];
      cursor += sizes[i];
    }
    return true;
  }

 private:
  std::unique_ptr<CostFunctor> functor_;
};

// ----------------------------------------------------------------------- problem
namespace internal {
class ProblemImpl;
}

class Problem {
 public:
  Problem();
  ~Problem();
  Problem(const Problem&) = delete;
  Problem& operator=(const Problem&) = delete;

  // The problem takes ownership of cost_function (upstream default TAKE_OWNERSHIP).
  // loss_function must be NULL (the reference passes NULL, sfm.cc:48,92).
  void* AddResidualBlock(CostFunction* cost_function, LossFunction* loss_function,
                         const std::vector<double*>& parameter_blocks);
  template <typename... Ts>
  void* AddResidualBlock(CostFunction* cost_function, LossFunction* loss_function, double* x0,
                         Ts*... xs) {
    const std::vector<double*> blocks({x0, xs...});
    return AddResidualBlock(cost_function, loss_function, blocks);
  }
  void AddParameterBlock(double* values, int size);
  void SetParameterBlockConstant(double* values);
  void SetParameterBlockVariable(double* values);
  // (shim extension, not upstream) blocks with the same id >= 0 form one diagonal block of the
  // block-Jacobi preconditioner of the implicit-Schur PCG mode; see mini_ceres.cc::SolveImplicit.
  void SetParameterBlockPreconditionerGroup(double* values, int group);
  int NumResidualBlocks() const;
  int NumParameterBlocks() const;

  internal::ProblemImpl* impl() { return impl_; }

 private:
  internal::ProblemImpl* impl_;
};

// ------------------------------------------------------------------------ solver
enum LinearSolverType {
  DENSE_NORMAL_CHOLESKY,
  DENSE_QR,
  SPARSE_NORMAL_CHOLESKY,
  DENSE_SCHUR,
  SPARSE_SCHUR,
  ITERATIVE_SCHUR,
  CGNR
};
enum PreconditionerType { IDENTITY, JACOBI, SCHUR_JACOBI };
enum TerminationType { CONVERGENCE, NO_CONVERGENCE, FAILURE, USER_SUCCESS, USER_FAILURE };

struct IterationSummary {
  int iteration = 0;
  bool step_is_valid = false;
  bool step_is_successful = false;
  double cost = 0.0;
  double cost_change = 0.0;
  double gradient_max_norm = 0.0;
  double gradient_norm = 0.0;
  double step_norm = 0.0;
  double relative_decrease = 0.0;
  double trust_region_radius = 0.0;
  double eta = 0.0;
  int linear_solver_iterations = 0;
  double iteration_time_in_seconds = 0.0;
  double cumulative_time_in_seconds = 0.0;
  // shim extension: model_cost_change of the step (for trace parity tests)
  double model_cost_change = 0.0;
};

class Solver {
 public:
  struct Options {
    // defaults == upstream defaults; the reference overrides only the five
    // marked fields (sfm.cc:66-71, 94-99).
    LinearSolverType linear_solver_type = SPARSE_NORMAL_CHOLESKY;  // overridden: DENSE_SCHUR
    PreconditionerType preconditioner_type = JACOBI;
    bool minimizer_progress_to_stdout = false;  // overridden: true
    int max_num_iterations = 50;                // overridden: 100 / 1000
    int num_threads = 1;                        // overridden: 16 / 20
    double max_solver_time_in_seconds = 1e9;    // overridden: 3600
    double initial_trust_region_radius = 1e4;
    double max_trust_region_radius = 1e16;
    double min_trust_region_radius = 1e-32;
    double min_relative_decrease = 1e-3;
    double min_lm_diagonal = 1e-6;
    double max_lm_diagonal = 1e32;
    int max_num_consecutive_invalid_steps = 5;
    double function_tolerance = 1e-6;
    double gradient_tolerance = 1e-10;
    double parameter_tolerance = 1e-8;
    bool jacobi_scaling = true;
    int max_linear_solver_iterations = 500;
    int min_linear_solver_iterations = 0;
    double eta = 1e-1;
    // ---- shim extensions (used by the oracle's implicit-Schur mode only) ----
    // ITERATIVE_SCHUR here = implicit Schur complement + block-Jacobi(S) PCG with
    // the SAME stopping rule as the GPU engine: stop when ||r_k|| <= shim_pcg_rel_tol*||r_0||
    // or after max_linear_solver_iterations; never fewer than min_linear_solver_iterations.
    double shim_pcg_rel_tol = 1e-12;
  };

  struct Summary {
    std::string message;
    TerminationType termination_type = NO_CONVERGENCE;
    double initial_cost = 0.0;
    double final_cost = 0.0;
    double fixed_cost = 0.0;
    std::vector<IterationSummary> iterations;
    int num_successful_steps = 0;
    int num_unsuccessful_steps = 0;
    double total_time_in_seconds = 0.0;
    double residual_evaluation_time_in_seconds = 0.0;
    double jacobian_evaluation_time_in_seconds = 0.0;
    double linear_solver_time_in_seconds = 0.0;
    int num_residual_evaluations = 0;
    int num_jacobian_evaluations = 0;
    int num_linear_solves = 0;
    int num_parameter_blocks = 0, num_parameters = 0, num_residual_blocks = 0, num_residuals = 0;
    int num_parameter_blocks_reduced = 0, num_parameters_reduced = 0;
    int num_residual_blocks_reduced = 0, num_residuals_reduced = 0;
    int num_e_blocks = 0, num_f_blocks = 0, reduced_system_size = 0;
    int num_threads_given = 0, num_threads_used = 0;
    LinearSolverType linear_solver_type_used = DENSE_SCHUR;
    std::string BriefReport() const;
    std::string FullReport() const;
  };
};

void Solve(const Solver::Options& options, Problem* problem, Solver::Summary* summary);

namespace shim {
// The reference's solve() (sfm.cc:31-75) discards its Summary after printing it;
// the shim keeps a copy of the most recent one so the bridge can read the trace.
const Solver::Summary& LastSummary();
// Overrides applied on top of whatever Options the caller passes to Solve() —
// lets tests drive the UNMODIFIED reference solve() with e.g. zero tolerances,
// fewer threads or quiet output.  Negative / NaN fields mean "leave as given".
struct Overrides {
  int quiet = -1;                    // 1: force minimizer_progress_to_stdout=false
  int num_threads = -1;
  int max_num_iterations = -1;
  double function_tolerance = -1.0;
  double gradient_tolerance = -1.0;
  double parameter_tolerance = -1.0;
  int linear_solver_type = -1;       // cast of LinearSolverType
  int max_linear_solver_iterations = -1;
  int min_linear_solver_iterations = -1;
  double shim_pcg_rel_tol = -1.0;
};
Overrides& GlobalOverrides();
}  // namespace shim

}  // namespace ceres

#endif  // ORACLE_CERES_SHIM_CERES_H_
