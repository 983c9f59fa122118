// TEST INFRASTRUCTURE — part of the CPU oracle, never linked into the product library.
//
// mini-Ceres solver: CPU restatement of what `ceres::Solve` does for the options
// the reference sets (sfm.cc:66-73, 94-101): trust-region Levenberg-Marquardt with
// Jacobi column scaling and an exact DENSE_SCHUR linear solve.  Every block below
// restates upstream ceres-solver 2.x behaviour [Ceres-upstream] and names the
// upstream unit it follows; none of it can be checked against a real Ceres in this
// image (PARITY UNPINNED, see ceres.h).
//
//   Preprocess()      trust_region_preprocessor.cc / reduced program, fixed cost,
//                     stable independent-set Schur ordering
//   Evaluate()        program_evaluator.h (cost = 1/2 sum r^2, g = J^T r)
//   SolveDenseSchur() schur_eliminator_impl.h + dense Cholesky of the reduced system
//   SolveImplicit()   (shim extension) implicit Schur complement + block-Jacobi PCG,
//                     same stopping rule as the GPU engine — used only as the CPU
//                     baseline where the dense reduced matrix does not fit
//   Solve()           trust_region_minimizer.cc + levenberg_marquardt_strategy.cc
#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <sstream>
#include <unordered_map>

#include "ceres/ceres.h"

namespace ceres {
namespace internal {

struct ParamRecord {
  double* user;
  int size;
  bool constant;
  int group = -1;  // shim extension: preconditioner group of the implicit-Schur PCG (-1: the block alone)
};

struct ResidualRecord {
  CostFunction* cost;
  std::vector<double*> params;
  LossFunction* loss = nullptr;
};

class ProblemImpl {
 public:
  std::vector<ParamRecord> params;
  std::unordered_map<double*, int> index_of;
  std::vector<ResidualRecord> residuals;
  std::set<CostFunction*> owned;
  std::set<LossFunction*> owned_loss;

  int Intern(double* p, int size) {
    auto it = index_of.find(p);
    if (it != index_of.end()) {
      if (params[it->second].size != size) {
        std::fprintf(stderr, "mini-ceres: parameter block %p re-added with size %d != %d\n",
                     static_cast<void*>(p), size, params[it->second].size);
        std::abort();
      }
      return it->second;
    }
    const int id = static_cast<int>(params.size());
    params.push_back(ParamRecord{p, size, false});
    index_of.emplace(p, id);
    return id;
  }
};

}  // namespace internal

Problem::Problem() : impl_(new internal::ProblemImpl) {}
Problem::~Problem() {
  for (CostFunction* c : impl_->owned) delete c;
  for (LossFunction* l : impl_->owned_loss) delete l;
  delete impl_;
}

void* Problem::AddResidualBlock(CostFunction* cost_function, LossFunction* loss_function,
                                const std::vector<double*>& parameter_blocks) {
  const std::vector<int32_t>& sizes = cost_function->parameter_block_sizes();
  if (sizes.size() != parameter_blocks.size()) {
    std::fprintf(stderr, "mini-ceres: cost function expects %zu parameter blocks, got %zu\n",
                 sizes.size(), parameter_blocks.size());
    std::abort();
  }
  for (size_t i = 0; i < sizes.size(); ++i) impl_->Intern(parameter_blocks[i], sizes[i]);
  impl_->residuals.push_back(internal::ResidualRecord{cost_function, parameter_blocks, loss_function});
  impl_->owned.insert(cost_function);
  if (loss_function != NULL) impl_->owned_loss.insert(loss_function);  // TAKE_OWNERSHIP, possibly shared
  return &impl_->residuals.back();
}

void Problem::AddParameterBlock(double* values, int size) { impl_->Intern(values, size); }

void Problem::SetParameterBlockConstant(double* values) {
  auto it = impl_->index_of.find(values);
  if (it == impl_->index_of.end()) {
    std::fprintf(stderr, "mini-ceres: SetParameterBlockConstant on unknown block\n");
    std::abort();
  }
  impl_->params[it->second].constant = true;
}
// (shim extension) parameter blocks with the same group id >= 0 form ONE diagonal block of the
// block-Jacobi preconditioner of the implicit-Schur PCG (SolveImplicit); no effect on DENSE_SCHUR.
void Problem::SetParameterBlockPreconditionerGroup(double* values, int group) {
  auto it = impl_->index_of.find(values);
  if (it == impl_->index_of.end()) {
    std::fprintf(stderr, "mini-ceres: SetParameterBlockPreconditionerGroup on an unknown block\n");
    std::abort();
  }
  impl_->params[it->second].group = group;
}

void Problem::SetParameterBlockVariable(double* values) {
  auto it = impl_->index_of.find(values);
  if (it != impl_->index_of.end()) impl_->params[it->second].constant = false;
}
int Problem::NumResidualBlocks() const { return static_cast<int>(impl_->residuals.size()); }
int Problem::NumParameterBlocks() const { return static_cast<int>(impl_->params.size()); }

namespace {

double Now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch())
      .count();
}

// ---- reduced program ----------------------------------------------------------
struct FreeBlock {
  double* user;
  int group = -1;  // preconditioner group (shim extension)
  int size;
  int offset;     // offset in x (problem order)
  bool is_e;      // eliminated in the Schur complement
  int ef_offset;  // offset within e-space (is_e) or within the reduced (f) system
  int ef_index;   // index among e blocks or among f blocks
};

struct RBlock {
  const CostFunction* cost;
  const LossFunction* loss = nullptr;
  int nres;
  int row;                     // first residual row
  std::vector<double*> user;   // user pointers, one per cost-function slot
  std::vector<int> free_id;    // FreeBlock index per slot, -1 if constant
  std::vector<size_t> jac_off; // offset into jac storage per slot (free slots only)
  int e_slot;                  // slot holding the e-block, -1 if none
};

struct Program {
  std::vector<FreeBlock> blocks;
  std::vector<RBlock> rbs;
  int num_params = 0;
  int num_residuals = 0;
  int num_e = 0, num_f = 0;
  int e_size = 0, f_size = 0;  // scalar sizes
  double fixed_cost = 0.0;
  size_t jac_size = 0;
  // e-block -> residual blocks (CSR), plus the residual blocks without an e-block
  std::vector<int> e_first, e_rbs, no_e_rbs;
  std::vector<int> e_block_ids, f_block_ids;
  int max_block = 1, max_nres = 1, max_slots = 1;
};

void Preprocess(internal::ProblemImpl* pi, bool want_schur, Program* prog) {
  std::vector<int> free_of(pi->params.size(), -1);
  // parameter blocks keep problem order (order of first appearance)
  std::vector<char> used(pi->params.size(), 0);
  for (const auto& rr : pi->residuals)
    for (double* p : rr.params) used[pi->index_of[p]] = 1;
  int offset = 0;
  for (size_t i = 0; i < pi->params.size(); ++i) {
    const auto& pr = pi->params[i];
    if (!used[i] || pr.constant || pr.size == 0) continue;
    free_of[i] = static_cast<int>(prog->blocks.size());
    prog->blocks.push_back(FreeBlock{pr.user, pr.group, pr.size, offset, false, 0, 0});
    offset += pr.size;
    prog->max_block = std::max(prog->max_block, pr.size);
  }
  prog->num_params = offset;

  // residual blocks; those without any free parameter contribute to fixed_cost
  int row = 0;
  size_t joff = 0;
  for (const auto& rr : pi->residuals) {
    RBlock rb;
    rb.cost = rr.cost;
    rb.loss = rr.loss;
    rb.nres = rr.cost->num_residuals();
    rb.user = rr.params;
    rb.e_slot = -1;
    bool any_free = false;
    const auto& sizes = rr.cost->parameter_block_sizes();
    for (size_t s = 0; s < rr.params.size(); ++s) {
      const int fid = free_of[pi->index_of[rr.params[s]]];
      rb.free_id.push_back(fid);
      any_free |= fid >= 0;
    }
    if (!any_free) {
      std::vector<double> r(rb.nres);
      rr.cost->Evaluate(rr.params.data(), r.data(), NULL);
      double c = 0.0;
      for (double v : r) c += v * v;
      if (rr.loss != NULL) {
        double rho[3];
        rr.loss->Evaluate(c, rho);
        c = rho[0];
      }
      prog->fixed_cost += 0.5 * c;
      continue;
    }
    rb.row = row;
    row += rb.nres;
    rb.jac_off.assign(rr.params.size(), 0);
    for (size_t s = 0; s < rr.params.size(); ++s) {
      if (rb.free_id[s] < 0) continue;
      rb.jac_off[s] = joff;
      joff += static_cast<size_t>(rb.nres) * sizes[s];
    }
    prog->max_nres = std::max(prog->max_nres, rb.nres);
    prog->max_slots = std::max(prog->max_slots, static_cast<int>(rr.params.size()));
    prog->rbs.push_back(std::move(rb));
  }
  prog->num_residuals = row;
  prog->jac_size = joff;

  const int nb = static_cast<int>(prog->blocks.size());
  if (want_schur && nb > 0) {
    // Hessian graph: an edge between every two free blocks sharing a residual block
    // [Ceres-upstream parameter_block_ordering.cc CreateHessianGraph].
    std::vector<std::pair<int, int>> edges;
    for (const auto& rb : prog->rbs)
      for (size_t a = 0; a < rb.free_id.size(); ++a)
        for (size_t b = 0; b < rb.free_id.size(); ++b)
          if (a != b && rb.free_id[a] >= 0 && rb.free_id[b] >= 0 &&
              rb.free_id[a] != rb.free_id[b])
            edges.emplace_back(rb.free_id[a], rb.free_id[b]);
    std::sort(edges.begin(), edges.end());
    edges.erase(std::unique(edges.begin(), edges.end()), edges.end());
    std::vector<int> first(nb + 1, 0);
    for (const auto& e : edges) first[e.first + 1]++;
    for (int i = 0; i < nb; ++i) first[i + 1] += first[i];
    // StableIndependentSetOrdering: vertices by ascending degree (stable), greedy.
    std::vector<int> order(nb);
    for (int i = 0; i < nb; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
      return (first[a + 1] - first[a]) < (first[b + 1] - first[b]);
    });
    std::vector<char> colour(nb, 0);  // 0 white, 1 grey, 2 black
    for (int v : order) {
      if (colour[v] != 0) continue;
      colour[v] = 2;
      for (int k = first[v]; k < first[v + 1]; ++k)
        if (colour[edges[k].second] == 0) colour[edges[k].second] = 1;
    }
    for (int i = 0; i < nb; ++i) prog->blocks[i].is_e = colour[i] == 2;
  }
  int eo = 0, fo = 0;
  for (int i = 0; i < nb; ++i) {
    FreeBlock& b = prog->blocks[i];
    if (b.is_e) {
      b.ef_offset = eo;
      b.ef_index = prog->num_e++;
      eo += b.size;
      prog->e_block_ids.push_back(i);
    } else {
      b.ef_offset = fo;
      b.ef_index = prog->num_f++;
      fo += b.size;
      prog->f_block_ids.push_back(i);
    }
  }
  prog->e_size = eo;
  prog->f_size = fo;
  // chunks
  std::vector<int> count(prog->num_e + 1, 0);
  for (size_t r = 0; r < prog->rbs.size(); ++r) {
    RBlock& rb = prog->rbs[r];
    for (size_t s = 0; s < rb.free_id.size(); ++s)
      if (rb.free_id[s] >= 0 && prog->blocks[rb.free_id[s]].is_e) rb.e_slot = static_cast<int>(s);
    if (rb.e_slot >= 0)
      count[prog->blocks[rb.free_id[rb.e_slot]].ef_index + 1]++;
    else
      prog->no_e_rbs.push_back(static_cast<int>(r));
  }
  prog->e_first.assign(prog->num_e + 1, 0);
  for (int i = 0; i < prog->num_e; ++i) prog->e_first[i + 1] = prog->e_first[i] + count[i + 1];
  prog->e_rbs.resize(prog->e_first[prog->num_e]);
  std::vector<int> cur(prog->e_first.begin(), prog->e_first.end() - 1);
  for (size_t r = 0; r < prog->rbs.size(); ++r) {
    const RBlock& rb = prog->rbs[r];
    if (rb.e_slot < 0) continue;
    prog->e_rbs[cur[prog->blocks[rb.free_id[rb.e_slot]].ef_index]++] = static_cast<int>(r);
  }
}

// ---- evaluation ---------------------------------------------------------------
struct EvalState {
  std::vector<double> residuals, jac, gradient;
};

bool Evaluate(const Program& prog, const double* x, bool want_jac, double* cost,
              std::vector<double>* residuals, std::vector<double>* jac,
              std::vector<double>* gradient, int num_threads) {
  const int nrb = static_cast<int>(prog.rbs.size());
  residuals->resize(prog.num_residuals);
  if (want_jac) jac->resize(prog.jac_size);
  bool ok = true;
  double total = 0.0;
#pragma omp parallel num_threads(num_threads) reduction(+ : total)
  {
    std::vector<const double*> pp(prog.max_slots);
    std::vector<double*> jj(prog.max_slots);
#pragma omp for schedule(static)
    for (int r = 0; r < nrb; ++r) {
      const RBlock& rb = prog.rbs[r];
      const int ns = static_cast<int>(rb.user.size());
      for (int s = 0; s < ns; ++s) {
        const int fid = rb.free_id[s];
        pp[s] = fid >= 0 ? x + prog.blocks[fid].offset : rb.user[s];
        jj[s] = (want_jac && fid >= 0) ? jac->data() + rb.jac_off[s] : NULL;
      }
      double* res = residuals->data() + rb.row;
      if (!rb.cost->Evaluate(pp.data(), res, want_jac ? jj.data() : NULL)) {
#pragma omp atomic write
        ok = false;
      }
      double sq = 0.0;
      for (int k = 0; k < rb.nres; ++k) sq += res[k] * res[k];
      if (rb.loss == NULL) {
        total += sq;
        continue;
      }
      // robust loss [Ceres-upstream residual_block.cc + corrector.cc]: cost 1/2 rho(s); the residuals
      // and Jacobians handed to the linear solver are corrected so that J~^T J~ is the Gauss-Newton
      // (Triggs) approximation of the robustified Hessian
      double rho[3];
      rb.loss->Evaluate(sq, rho);
      total += rho[0];
      const double sqrt_rho1 = std::sqrt(rho[1]);
      double residual_scaling, alpha_sq_norm;
      if (sq == 0.0 || rho[2] <= 0.0) {
        residual_scaling = sqrt_rho1;
        alpha_sq_norm = 0.0;
      } else {
        const double D = 1.0 + 2.0 * sq * rho[2] / rho[1];
        const double alpha = 1.0 - std::sqrt(D);
        residual_scaling = sqrt_rho1 / (1 - alpha);
        alpha_sq_norm = alpha / sq;
      }
      if (want_jac) {
        for (int s = 0; s < ns; ++s) {
          if (jj[s] == NULL) continue;
          const int bs = prog.blocks[rb.free_id[s]].size;
          double* J = jj[s];
          if (alpha_sq_norm == 0.0) {
            for (int i = 0; i < rb.nres * bs; ++i) J[i] *= sqrt_rho1;
          } else {
            for (int c = 0; c < bs; ++c) {
              double r_dot_j = 0.0;
              for (int k = 0; k < rb.nres; ++k) r_dot_j += J[k * bs + c] * res[k];
              for (int k = 0; k < rb.nres; ++k) J[k * bs + c] = sqrt_rho1 * (J[k * bs + c] - alpha_sq_norm * res[k] * r_dot_j);
            }
          }
        }
      }
      for (int k = 0; k < rb.nres; ++k) res[k] *= residual_scaling;
    }
  }
  *cost = 0.5 * total;
  if (!std::isfinite(*cost)) ok = false;
  if (want_jac && gradient != NULL) {
    // g = J^T r; accumulate per free block through per-thread partial vectors.
    gradient->assign(prog.num_params, 0.0);
    const int nt = std::max(1, num_threads);
    std::vector<std::vector<double>> part(nt);
#pragma omp parallel num_threads(nt)
    {
      std::vector<double>& g = part[omp_get_thread_num()];
      g.assign(prog.num_params, 0.0);
#pragma omp for schedule(static)
      for (int r = 0; r < nrb; ++r) {
        const RBlock& rb = prog.rbs[r];
        const double* res = residuals->data() + rb.row;
        for (size_t s = 0; s < rb.user.size(); ++s) {
          const int fid = rb.free_id[s];
          if (fid < 0) continue;
          const FreeBlock& b = prog.blocks[fid];
          const double* J = jac->data() + rb.jac_off[s];
          for (int k = 0; k < rb.nres; ++k)
            for (int c = 0; c < b.size; ++c) g[b.offset + c] += J[k * b.size + c] * res[k];
        }
      }
    }
    for (int t = 0; t < nt; ++t)
      if (!part[t].empty())
        for (int i = 0; i < prog.num_params; ++i) (*gradient)[i] += part[t][i];
  }
  return ok;
}

void SquaredColumnNorms(const Program& prog, const std::vector<double>& jac,
                        std::vector<double>* out) {
  out->assign(prog.num_params, 0.0);
  for (const RBlock& rb : prog.rbs)
    for (size_t s = 0; s < rb.user.size(); ++s) {
      const int fid = rb.free_id[s];
      if (fid < 0) continue;
      const FreeBlock& b = prog.blocks[fid];
      const double* J = jac.data() + rb.jac_off[s];
      for (int k = 0; k < rb.nres; ++k)
        for (int c = 0; c < b.size; ++c) (*out)[b.offset + c] += J[k * b.size + c] * J[k * b.size + c];
    }
}

void ScaleColumns(const Program& prog, const std::vector<double>& scale, std::vector<double>* jac,
                  int num_threads) {
  const int nrb = static_cast<int>(prog.rbs.size());
#pragma omp parallel for schedule(static) num_threads(num_threads)
  for (int r = 0; r < nrb; ++r) {
    const RBlock& rb = prog.rbs[r];
    for (size_t s = 0; s < rb.user.size(); ++s) {
      const int fid = rb.free_id[s];
      if (fid < 0) continue;
      const FreeBlock& b = prog.blocks[fid];
      double* J = jac->data() + rb.jac_off[s];
      for (int k = 0; k < rb.nres; ++k)
        for (int c = 0; c < b.size; ++c) J[k * b.size + c] *= scale[b.offset + c];
    }
  }
}

// y = J * v (v in x-order), used for model_cost_change
void RightMultiply(const Program& prog, const std::vector<double>& jac, const double* v,
                   std::vector<double>* y, int num_threads) {
  y->assign(prog.num_residuals, 0.0);
  const int nrb = static_cast<int>(prog.rbs.size());
#pragma omp parallel for schedule(static) num_threads(num_threads)
  for (int r = 0; r < nrb; ++r) {
    const RBlock& rb = prog.rbs[r];
    for (size_t s = 0; s < rb.user.size(); ++s) {
      const int fid = rb.free_id[s];
      if (fid < 0) continue;
      const FreeBlock& b = prog.blocks[fid];
      const double* J = jac.data() + rb.jac_off[s];
      for (int k = 0; k < rb.nres; ++k) {
        double acc = 0.0;
        for (int c = 0; c < b.size; ++c) acc += J[k * b.size + c] * v[b.offset + c];
        (*y)[rb.row + k] += acc;
      }
    }
  }
}

// ---- small dense helpers ------------------------------------------------------
// In-place lower Cholesky of a row-major n x n SPD matrix; returns false if not SPD.
bool CholeskySmall(double* A, int n) {
  for (int j = 0; j < n; ++j) {
    double d = A[j * n + j];
    for (int k = 0; k < j; ++k) d -= A[j * n + k] * A[j * n + k];
    if (!(d > 0.0) || !std::isfinite(d)) return false;
    d = std::sqrt(d);
    A[j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[i * n + j];
      for (int k = 0; k < j; ++k) s -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = s / d;
    }
  }
  return true;
}

// inv = A^{-1} for SPD A (row-major n x n) via Cholesky; A is destroyed.
bool InvertSPD(double* A, int n, double* inv) {
  if (!CholeskySmall(A, n)) return false;
  for (int c = 0; c < n; ++c) {
    // solve L L^T x = e_c
    std::vector<double> y(n, 0.0);
    for (int i = 0; i < n; ++i) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int k = 0; k < i; ++k) s -= A[i * n + k] * y[k];
      y[i] = s / A[i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = y[i];
      for (int k = i + 1; k < n; ++k) s -= A[k * n + i] * inv[k * n + c];
      inv[i * n + c] = s / A[i * n + i];
    }
  }
  return true;
}

// Blocked, OpenMP-parallel in-place lower Cholesky (row-major) for the reduced system.
bool CholeskyBlocked(double* A, int n, int num_threads) {
  const int B = 96;
  if (n <= 2 * B) return CholeskySmall(A, n);
  bool ok = true;
  for (int k0 = 0; k0 < n && ok; k0 += B) {
    const int kb = std::min(B, n - k0);
    // factor the diagonal block
    for (int j = k0; j < k0 + kb; ++j) {
      double d = A[(size_t)j * n + j];
      for (int k = k0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
      if (!(d > 0.0) || !std::isfinite(d)) {
        ok = false;
        break;
      }
      d = std::sqrt(d);
      A[(size_t)j * n + j] = d;
      for (int i = j + 1; i < k0 + kb; ++i) {
        double s = A[(size_t)i * n + j];
        for (int k = k0; k < j; ++k) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
        A[(size_t)i * n + j] = s / d;
      }
    }
    if (!ok) break;
    const int rest = n - (k0 + kb);
    if (rest <= 0) break;
    // panel: rows below solve  X * L_kk^T = A_ik
#pragma omp parallel for schedule(static) num_threads(num_threads)
    for (int i = k0 + kb; i < n; ++i) {
      double* Ai = A + (size_t)i * n;
      for (int j = k0; j < k0 + kb; ++j) {
        const double* Aj = A + (size_t)j * n;
        double s = Ai[j];
        for (int k = k0; k < j; ++k) s -= Ai[k] * Aj[k];
        Ai[j] = s / Aj[j];
      }
    }
    // trailing update A_ij -= L_ik L_jk^T for j <= i, tiled
    const int nt = (rest + B - 1) / B;
#pragma omp parallel for schedule(dynamic) collapse(2) num_threads(num_threads)
    for (int ti = 0; ti < nt; ++ti)
      for (int tj = 0; tj < nt; ++tj) {
        if (tj > ti) continue;
        const int i0 = k0 + kb + ti * B, i1 = std::min(n, i0 + B);
        const int j0 = k0 + kb + tj * B, j1 = std::min(n, j0 + B);
        for (int i = i0; i < i1; ++i) {
          const double* Li = A + (size_t)i * n + k0;
          double* Ai = A + (size_t)i * n;
          const int jend = std::min(j1, i + 1);
          for (int j = j0; j < jend; ++j) {
            const double* Lj = A + (size_t)j * n + k0;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int k = 0;
            for (; k + 4 <= kb; k += 4) {
              s0 += Li[k] * Lj[k];
              s1 += Li[k + 1] * Lj[k + 1];
              s2 += Li[k + 2] * Lj[k + 2];
              s3 += Li[k + 3] * Lj[k + 3];
            }
            for (; k < kb; ++k) s0 += Li[k] * Lj[k];
            Ai[j] -= (s0 + s1) + (s2 + s3);
          }
        }
      }
  }
  return ok;
}

void CholeskySolveInPlace(const double* L, int n, double* b) {
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    const double* Li = L + (size_t)i * n;
    for (int k = 0; k < i; ++k) s -= Li[k] * b[k];
    b[i] = s / Li[i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int k = i + 1; k < n; ++k) s -= L[(size_t)k * n + i] * b[k];
    b[i] = s / L[(size_t)i * n + i];
  }
}

// ---- Schur machinery ----------------------------------------------------------
// Solves (J^T J + D^2) y = J^T r exactly:  eliminate the e-blocks, dense Cholesky
// on the reduced system, back-substitute [Ceres-upstream schur_eliminator_impl.h,
// schur_complement_solver.cc DenseSchurComplementSolver].  y is in x-order.
struct SchurWork {
  std::vector<double> ete_inv;     // per e-block, size^2 each (offsets below)
  std::vector<size_t> ete_off;
  std::vector<double> S, rhs;
};

// per-chunk scratch
struct ChunkScratch {
  std::vector<int> slot_of;          // f-block index -> local slot or -1
  std::vector<int> local_f;          // local slot -> f ef_index
  std::vector<double> FtE;           // local slot -> (fsize x esize), packed at stride max_block^2
  std::vector<double> ete, g, tmp, tmp2;
};

bool EliminateAndSolve(const Program& prog, const std::vector<double>& jac,
                       const std::vector<double>& residuals, const std::vector<double>& D,
                       int num_threads, double* y, SchurWork* w) {
  const int nf = prog.f_size;
  const int ne_blocks = prog.num_e;
  const int mb = prog.max_block;
  w->ete_off.resize(ne_blocks + 1);
  size_t eo = 0;
  for (int e = 0; e < ne_blocks; ++e) {
    w->ete_off[e] = eo;
    const int s = prog.blocks[prog.e_block_ids[e]].size;
    eo += static_cast<size_t>(s) * s;
  }
  w->ete_off[ne_blocks] = eo;
  w->ete_inv.assign(eo, 0.0);
  w->S.assign(static_cast<size_t>(nf) * nf, 0.0);
  w->rhs.assign(nf, 0.0);

  const int nt = std::max(1, num_threads);
  // Small reduced systems: thread-private copies reduced at the end; large: atomics.
  const bool private_S = static_cast<double>(nf) * nf * nt * 8.0 <= 512e6;
  std::vector<std::vector<double>> Sp(private_S ? nt : 0), rp(private_S ? nt : 0);
  bool ok = true;

#pragma omp parallel num_threads(nt)
  {
    const int tid = omp_get_thread_num();
    double* S;
    double* rhs;
    if (private_S) {
      Sp[tid].assign(static_cast<size_t>(nf) * nf, 0.0);
      rp[tid].assign(nf, 0.0);
      S = Sp[tid].data();
      rhs = rp[tid].data();
    } else {
      S = w->S.data();
      rhs = w->rhs.data();
    }
    auto add = [&](double* dst, double v) {
      if (private_S) {
        *dst += v;
      } else {
#pragma omp atomic
        *dst += v;
      }
    };
    ChunkScratch cs;
    cs.slot_of.assign(prog.num_f, -1);
    cs.ete.resize(mb * mb);
    cs.g.resize(mb);
    cs.tmp.resize(mb * mb);
    cs.tmp2.resize(mb * mb);

    // F^T F and F^T r contributions of one residual block (upper block triangle)
    auto add_ftf = [&](const RBlock& rb) {
      const double* res = residuals.data() + rb.row;
      for (size_t a = 0; a < rb.user.size(); ++a) {
        const int fa = rb.free_id[a];
        if (fa < 0 || prog.blocks[fa].is_e) continue;
        const FreeBlock& ba = prog.blocks[fa];
        const double* Ja = jac.data() + rb.jac_off[a];
        for (int c = 0; c < ba.size; ++c) {
          double acc = 0.0;
          for (int k = 0; k < rb.nres; ++k) acc += Ja[k * ba.size + c] * res[k];
          add(&rhs[ba.ef_offset + c], acc);
        }
        for (size_t b = 0; b < rb.user.size(); ++b) {
          const int fb = rb.free_id[b];
          if (fb < 0 || prog.blocks[fb].is_e) continue;
          const FreeBlock& bb = prog.blocks[fb];
          if (bb.ef_offset < ba.ef_offset) continue;
          const double* Jb = jac.data() + rb.jac_off[b];
          for (int i = 0; i < ba.size; ++i)
            for (int j = 0; j < bb.size; ++j) {
              double acc = 0.0;
              for (int k = 0; k < rb.nres; ++k) acc += Ja[k * ba.size + i] * Jb[k * bb.size + j];
              add(&S[static_cast<size_t>(ba.ef_offset + i) * nf + bb.ef_offset + j], acc);
            }
        }
      }
    };

#pragma omp for schedule(dynamic, 64)
    for (int e = 0; e < ne_blocks; ++e) {
      const FreeBlock& eb = prog.blocks[prog.e_block_ids[e]];
      const int es = eb.size;
      double* ete = cs.ete.data();
      double* g = cs.g.data();
      for (int i = 0; i < es * es; ++i) ete[i] = 0.0;
      for (int i = 0; i < es; ++i) {
        ete[i * es + i] = D[eb.offset + i] * D[eb.offset + i];
        g[i] = 0.0;
      }
      cs.local_f.clear();
      for (int q = prog.e_first[e]; q < prog.e_first[e + 1]; ++q) {
        const RBlock& rb = prog.rbs[prog.e_rbs[q]];
        const double* E = jac.data() + rb.jac_off[rb.e_slot];
        const double* res = residuals.data() + rb.row;
        for (int i = 0; i < es; ++i) {
          for (int j = 0; j < es; ++j) {
            double acc = 0.0;
            for (int k = 0; k < rb.nres; ++k) acc += E[k * es + i] * E[k * es + j];
            ete[i * es + j] += acc;
          }
          double acc = 0.0;
          for (int k = 0; k < rb.nres; ++k) acc += E[k * es + i] * res[k];
          g[i] += acc;
        }
        // F^T E per distinct f-block of the chunk
        for (size_t s = 0; s < rb.user.size(); ++s) {
          const int fid = rb.free_id[s];
          if (fid < 0 || prog.blocks[fid].is_e) continue;
          const FreeBlock& fb = prog.blocks[fid];
          int slot = cs.slot_of[fb.ef_index];
          if (slot < 0) {
            slot = static_cast<int>(cs.local_f.size());
            cs.slot_of[fb.ef_index] = slot;
            cs.local_f.push_back(fid);
            if (cs.FtE.size() < static_cast<size_t>(slot + 1) * mb * mb)
              cs.FtE.resize(static_cast<size_t>(slot + 1) * mb * mb);
            std::fill(cs.FtE.begin() + static_cast<size_t>(slot) * mb * mb,
                      cs.FtE.begin() + static_cast<size_t>(slot + 1) * mb * mb, 0.0);
          }
          double* fte = cs.FtE.data() + static_cast<size_t>(slot) * mb * mb;
          const double* F = jac.data() + rb.jac_off[s];
          for (int i = 0; i < fb.size; ++i)
            for (int j = 0; j < es; ++j) {
              double acc = 0.0;
              for (int k = 0; k < rb.nres; ++k) acc += F[k * fb.size + i] * E[k * es + j];
              fte[i * es + j] += acc;
            }
        }
        add_ftf(rb);
      }
      double* inv = w->ete_inv.data() + w->ete_off[e];
      std::vector<double> ete_copy(ete, ete + es * es);
      if (!InvertSPD(ete_copy.data(), es, inv)) {
#pragma omp atomic write
        ok = false;
      }
      // tmp_g = ete^{-1} g
      double invg[16];
      std::vector<double> invg_dyn;
      double* ig = invg;
      if (es > 16) {
        invg_dyn.resize(es);
        ig = invg_dyn.data();
      }
      for (int i = 0; i < es; ++i) {
        double acc = 0.0;
        for (int j = 0; j < es; ++j) acc += inv[i * es + j] * g[j];
        ig[i] = acc;
      }
      const int nl = static_cast<int>(cs.local_f.size());
      for (int a = 0; a < nl; ++a) {
        const FreeBlock& fa = prog.blocks[cs.local_f[a]];
        const double* ftea = cs.FtE.data() + static_cast<size_t>(a) * mb * mb;
        // rhs_a -= FtE_a * ete^{-1} g
        for (int i = 0; i < fa.size; ++i) {
          double acc = 0.0;
          for (int j = 0; j < es; ++j) acc += ftea[i * es + j] * ig[j];
          add(&rhs[fa.ef_offset + i], -acc);
        }
        // tmp = FtE_a * ete^{-1}   (fa.size x es)
        double* tmp = cs.tmp.data();
        for (int i = 0; i < fa.size; ++i)
          for (int j = 0; j < es; ++j) {
            double acc = 0.0;
            for (int k = 0; k < es; ++k) acc += ftea[i * es + k] * inv[k * es + j];
            tmp[i * es + j] = acc;
          }
        for (int b = 0; b < nl; ++b) {
          const FreeBlock& fbk = prog.blocks[cs.local_f[b]];
          if (fbk.ef_offset < fa.ef_offset) continue;
          const double* fteb = cs.FtE.data() + static_cast<size_t>(b) * mb * mb;
          for (int i = 0; i < fa.size; ++i)
            for (int j = 0; j < fbk.size; ++j) {
              double acc = 0.0;
              for (int k = 0; k < es; ++k) acc += tmp[i * es + k] * fteb[j * es + k];
              add(&S[static_cast<size_t>(fa.ef_offset + i) * nf + fbk.ef_offset + j], -acc);
            }
        }
      }
      for (int a = 0; a < nl; ++a) cs.slot_of[prog.blocks[cs.local_f[a]].ef_index] = -1;
    }

    const int nno = static_cast<int>(prog.no_e_rbs.size());
#pragma omp for schedule(static)
    for (int q = 0; q < nno; ++q) add_ftf(prog.rbs[prog.no_e_rbs[q]]);
  }
  if (private_S) {
    const size_t n2 = static_cast<size_t>(nf) * nf;
#pragma omp parallel for schedule(static) num_threads(nt)
    for (size_t i = 0; i < n2; ++i) {
      double acc = 0.0;
      for (int t = 0; t < nt; ++t)
        if (!Sp[t].empty()) acc += Sp[t][i];
      w->S[i] = acc;
    }
    for (int t = 0; t < nt; ++t)
      if (!rp[t].empty())
        for (int i = 0; i < nf; ++i) w->rhs[i] += rp[t][i];
  }
  if (!ok) return false;
  // D_f^2 on the diagonal, mirror the upper triangle into the lower one
  for (int fb : prog.f_block_ids) {
    const FreeBlock& b = prog.blocks[fb];
    for (int i = 0; i < b.size; ++i)
      w->S[static_cast<size_t>(b.ef_offset + i) * nf + b.ef_offset + i] +=
          D[b.offset + i] * D[b.offset + i];
  }
#pragma omp parallel for schedule(static) num_threads(nt)
  for (int i = 0; i < nf; ++i)
    for (int j = 0; j < i; ++j)
      w->S[static_cast<size_t>(i) * nf + j] = w->S[static_cast<size_t>(j) * nf + i];
  // dense Cholesky of the reduced system
  if (nf > 0) {
    if (!CholeskyBlocked(w->S.data(), nf, nt)) return false;
    CholeskySolveInPlace(w->S.data(), nf, w->rhs.data());
  }
  for (int fb : prog.f_block_ids) {
    const FreeBlock& b = prog.blocks[fb];
    for (int i = 0; i < b.size; ++i) y[b.offset + i] = w->rhs[b.ef_offset + i];
  }
  // back-substitution: y_e = ete^{-1} (E^T r - E^T F y_f)
#pragma omp parallel for schedule(dynamic, 64) num_threads(nt)
  for (int e = 0; e < ne_blocks; ++e) {
    const FreeBlock& eb = prog.blocks[prog.e_block_ids[e]];
    const int es = eb.size;
    std::vector<double> acc(es, 0.0);
    for (int q = prog.e_first[e]; q < prog.e_first[e + 1]; ++q) {
      const RBlock& rb = prog.rbs[prog.e_rbs[q]];
      const double* E = jac.data() + rb.jac_off[rb.e_slot];
      const double* res = residuals.data() + rb.row;
      for (int k = 0; k < rb.nres; ++k) {
        double t = res[k];
        for (size_t s = 0; s < rb.user.size(); ++s) {
          const int fid = rb.free_id[s];
          if (fid < 0 || prog.blocks[fid].is_e) continue;
          const FreeBlock& fb = prog.blocks[fid];
          const double* F = jac.data() + rb.jac_off[s];
          for (int c = 0; c < fb.size; ++c) t -= F[k * fb.size + c] * y[fb.offset + c];
        }
        for (int i = 0; i < es; ++i) acc[i] += E[k * es + i] * t;
      }
    }
    const double* inv = w->ete_inv.data() + w->ete_off[e];
    for (int i = 0; i < es; ++i) {
      double s = 0.0;
      for (int j = 0; j < es; ++j) s += inv[i * es + j] * acc[j];
      y[eb.offset + i] = s;
    }
  }
  return true;
}

// (shim extension) implicit Schur complement + block-Jacobi(S) preconditioned CG.
// Same algorithm and stopping rule as the GPU engine (DESIGN.md "PCG"): solve
//   S y_f = b,  S = F^T F + D_f^2 - F^T E C^{-1} E^T F,  C = E^T E + D_e^2,
//   b = F^T r - F^T E C^{-1} E^T r,
// with M = blockdiag(S); stop when r_k^T z_k <= tol^2 r_0^T z_0 or k == max_iter.
bool SolveImplicit(const Program& prog, const std::vector<double>& jac,
                   const std::vector<double>& residuals, const std::vector<double>& D,
                   int num_threads, int min_iter, int max_iter, double rel_tol, double* y,
                   int* iterations) {
  const int nf = prog.f_size;
  const int ne_blocks = prog.num_e;
  const int nt = std::max(1, num_threads);
  const int nrb = static_cast<int>(prog.rbs.size());
  std::vector<size_t> ete_off(ne_blocks + 1, 0);
  for (int e = 0; e < ne_blocks; ++e) {
    const int s = prog.blocks[prog.e_block_ids[e]].size;
    ete_off[e + 1] = ete_off[e] + static_cast<size_t>(s) * s;
  }
  std::vector<double> ete_inv(ete_off[ne_blocks], 0.0);
  // block-diagonal of S, b.  One diagonal block per PRECONDITIONER GROUP: the f-blocks that carry the
  // same group id (the rotation / translation [/ focal / distortion] blocks of one camera when the
  // caller says so, as the GPU engine's 6- or 9-dof camera blocks) or, without a group, the
  // parameter block alone (what Ceres' SCHUR_JACOBI does with the reference's parameter blocks).
  std::vector<int> g_of(prog.num_f, -1), f_goff(prog.num_f, 0), g_size;
  std::vector<std::vector<int>> g_members;
  {
    std::unordered_map<int, int> seen;
    for (int f = 0; f < prog.num_f; ++f) {
      const FreeBlock& fb = prog.blocks[prog.f_block_ids[f]];
      int g;
      auto it = fb.group >= 0 ? seen.find(fb.group) : seen.end();
      if (it != seen.end()) {
        g = it->second;
      } else {
        g = static_cast<int>(g_size.size());
        g_size.push_back(0);
        g_members.emplace_back();
        if (fb.group >= 0) seen.emplace(fb.group, g);
      }
      g_of[f] = g;
      f_goff[f] = g_size[g];
      g_size[g] += fb.size;
      g_members[g].push_back(f);
    }
  }
  const int ng = static_cast<int>(g_size.size());
  std::vector<size_t> m_off(ng + 1, 0);
  for (int g = 0; g < ng; ++g) m_off[g + 1] = m_off[g] + static_cast<size_t>(g_size[g]) * g_size[g];
  std::vector<double> M(m_off[ng], 0.0), b(nf, 0.0);
  bool ok = true;
  std::vector<std::vector<double>> Mp(nt), bp(nt);
#pragma omp parallel num_threads(nt)
  {
    const int tid = omp_get_thread_num();
    Mp[tid].assign(M.size(), 0.0);
    bp[tid].assign(nf, 0.0);
    double* Ml = Mp[tid].data();
    double* bl = bp[tid].data();
    std::vector<double> ete, g, ig, fte, tmp;
#pragma omp for schedule(dynamic, 64)
    for (int e = 0; e < ne_blocks; ++e) {
      const FreeBlock& eb = prog.blocks[prog.e_block_ids[e]];
      const int es = eb.size;
      ete.assign(es * es, 0.0);
      g.assign(es, 0.0);
      ig.assign(es, 0.0);
      for (int i = 0; i < es; ++i) ete[i * es + i] = D[eb.offset + i] * D[eb.offset + i];
      for (int q = prog.e_first[e]; q < prog.e_first[e + 1]; ++q) {
        const RBlock& rb = prog.rbs[prog.e_rbs[q]];
        const double* E = jac.data() + rb.jac_off[rb.e_slot];
        const double* res = residuals.data() + rb.row;
        for (int i = 0; i < es; ++i) {
          for (int j = 0; j < es; ++j)
            for (int k = 0; k < rb.nres; ++k) ete[i * es + j] += E[k * es + i] * E[k * es + j];
          for (int k = 0; k < rb.nres; ++k) g[i] += E[k * es + i] * res[k];
        }
      }
      double* inv = ete_inv.data() + ete_off[e];
      if (!InvertSPD(ete.data(), es, inv)) {
#pragma omp atomic write
        ok = false;
      }
      for (int i = 0; i < es; ++i)
        for (int j = 0; j < es; ++j) ig[i] += inv[i * es + j] * g[j];
      // per residual block: M_ff += F^T F - (F^T E) C^{-1} (E^T F);  b_f += F^T (r - E C^{-1} g)
      // (the block-Jacobi term ignores cross terms between different residual blocks of
      //  the same (e,f) pair, exactly like the GPU engine: one observation per pair)
      for (int q = prog.e_first[e]; q < prog.e_first[e + 1]; ++q) {
        const RBlock& rb = prog.rbs[prog.e_rbs[q]];
        const double* E = jac.data() + rb.jac_off[rb.e_slot];
        const double* res = residuals.data() + rb.row;
        double rr[8];
        for (int k = 0; k < rb.nres && k < 8; ++k) {
          double t = res[k];
          for (int i = 0; i < es; ++i) t -= E[k * es + i] * ig[i];
          rr[k] = t;
        }
        // F^T E and (F^T E) C^-1 of every f slot of this residual block
        const int nslot = static_cast<int>(rb.user.size());
        const int mb2 = prog.max_block * es;
        fte.assign(static_cast<size_t>(nslot) * mb2, 0.0);
        tmp.assign(static_cast<size_t>(nslot) * mb2, 0.0);
        for (int s = 0; s < nslot; ++s) {
          const int fid = rb.free_id[s];
          if (fid < 0 || prog.blocks[fid].is_e) continue;
          const int fs = prog.blocks[fid].size;
          const double* F = jac.data() + rb.jac_off[s];
          double* ft = fte.data() + static_cast<size_t>(s) * mb2;
          double* tp = tmp.data() + static_cast<size_t>(s) * mb2;
          for (int i = 0; i < fs; ++i)
            for (int j = 0; j < es; ++j)
              for (int k = 0; k < rb.nres; ++k) ft[i * es + j] += F[k * fs + i] * E[k * es + j];
          for (int i = 0; i < fs; ++i)
            for (int j = 0; j < es; ++j)
              for (int k = 0; k < es; ++k) tp[i * es + j] += ft[i * es + k] * inv[k * es + j];
        }
        for (int s = 0; s < nslot; ++s) {
          const int fid = rb.free_id[s];
          if (fid < 0 || prog.blocks[fid].is_e) continue;
          const FreeBlock& fb = prog.blocks[fid];
          const double* F = jac.data() + rb.jac_off[s];
          const int fs = fb.size;
          const int g = g_of[fb.ef_index], gs = g_size[g], oi = f_goff[fb.ef_index];
          double* Mg = Ml + m_off[g];
          const double* tps = tmp.data() + static_cast<size_t>(s) * mb2;
          for (int s2 = 0; s2 < nslot; ++s2) {
            const int fid2 = rb.free_id[s2];
            if (fid2 < 0 || prog.blocks[fid2].is_e) continue;
            const FreeBlock& fb2 = prog.blocks[fid2];
            if (g_of[fb2.ef_index] != g) continue;
            const double* F2 = jac.data() + rb.jac_off[s2];
            const int fs2 = fb2.size, oj = f_goff[fb2.ef_index];
            const double* ft2 = fte.data() + static_cast<size_t>(s2) * mb2;
            for (int i = 0; i < fs; ++i)
              for (int j = 0; j < fs2; ++j) {
                double acc = 0.0;
                for (int k = 0; k < rb.nres; ++k) acc += F[k * fs + i] * F2[k * fs2 + j];
                for (int k = 0; k < es; ++k) acc -= tps[i * es + k] * ft2[j * es + k];
                Mg[(oi + i) * gs + oj + j] += acc;
              }
          }
          for (int i = 0; i < fs; ++i) {
            double acc = 0.0;
            for (int k = 0; k < rb.nres; ++k) acc += F[k * fs + i] * rr[k];
            bl[fb.ef_offset + i] += acc;
          }
        }
      }
    }
    const int nno = static_cast<int>(prog.no_e_rbs.size());
#pragma omp for schedule(static)
    for (int q = 0; q < nno; ++q) {
      const RBlock& rb = prog.rbs[prog.no_e_rbs[q]];
      const double* res = residuals.data() + rb.row;
      for (size_t s = 0; s < rb.user.size(); ++s) {
        const int fid = rb.free_id[s];
        if (fid < 0) continue;
        const FreeBlock& fb = prog.blocks[fid];
        const double* F = jac.data() + rb.jac_off[s];
        const int fs = fb.size;
        const int g = g_of[fb.ef_index], gs = g_size[g], oi = f_goff[fb.ef_index];
        double* Mg = Ml + m_off[g];
        for (size_t s2 = 0; s2 < rb.user.size(); ++s2) {
          const int fid2 = rb.free_id[s2];
          if (fid2 < 0 || g_of[prog.blocks[fid2].ef_index] != g) continue;
          const FreeBlock& fb2 = prog.blocks[fid2];
          const double* F2 = jac.data() + rb.jac_off[s2];
          const int fs2 = fb2.size, oj = f_goff[fb2.ef_index];
          for (int i = 0; i < fs; ++i)
            for (int j = 0; j < fs2; ++j)
              for (int k = 0; k < rb.nres; ++k) Mg[(oi + i) * gs + oj + j] += F[k * fs + i] * F2[k * fs2 + j];
        }
        for (int i = 0; i < fs; ++i)
          for (int k = 0; k < rb.nres; ++k) bl[fb.ef_offset + i] += F[k * fs + i] * res[k];
      }
    }
  }
  if (!ok) return false;
  for (int t = 0; t < nt; ++t) {
    for (size_t i = 0; i < M.size(); ++i) M[i] += Mp[t][i];
    for (int i = 0; i < nf; ++i) b[i] += bp[t][i];
  }
  // + D_f^2, invert the diagonal blocks (one per group)
  std::vector<double> Minv(M.size(), 0.0);
  for (int g = 0; g < ng; ++g) {
    double* Mg = M.data() + m_off[g];
    const int gs = g_size[g];
    for (int f : g_members[g]) {
      const FreeBlock& fb = prog.blocks[prog.f_block_ids[f]];
      for (int i = 0; i < fb.size; ++i) Mg[(f_goff[f] + i) * gs + f_goff[f] + i] += D[fb.offset + i] * D[fb.offset + i];
    }
    if (!InvertSPD(Mg, gs, Minv.data() + m_off[g])) return false;
  }
  auto apply_minv = [&](const std::vector<double>& r, std::vector<double>* z) {
    std::vector<double> rg, zg;
    for (int g = 0; g < ng; ++g) {
      const int gs = g_size[g];
      rg.assign(gs, 0.0);
      zg.assign(gs, 0.0);
      for (int f : g_members[g]) {
        const FreeBlock& fb = prog.blocks[prog.f_block_ids[f]];
        for (int i = 0; i < fb.size; ++i) rg[f_goff[f] + i] = r[fb.ef_offset + i];
      }
      const double* Mi = Minv.data() + m_off[g];
      for (int i = 0; i < gs; ++i) {
        double acc = 0.0;
        for (int j = 0; j < gs; ++j) acc += Mi[i * gs + j] * rg[j];
        zg[i] = acc;
      }
      for (int f : g_members[g]) {
        const FreeBlock& fb = prog.blocks[prog.f_block_ids[f]];
        for (int i = 0; i < fb.size; ++i) (*z)[fb.ef_offset + i] = zg[f_goff[f] + i];
      }
    }
  };
  // q = S p
  std::vector<std::vector<double>> qp(nt);
  auto apply_S = [&](const std::vector<double>& p, std::vector<double>* q) {
#pragma omp parallel num_threads(nt)
    {
      std::vector<double>& ql = qp[omp_get_thread_num()];
      ql.assign(nf, 0.0);
      std::vector<double> z, yv, u;
#pragma omp for schedule(dynamic, 64)
      for (int e = 0; e < ne_blocks; ++e) {
        const FreeBlock& eb = prog.blocks[prog.e_block_ids[e]];
        const int es = eb.size;
        z.assign(es, 0.0);
        yv.assign(es, 0.0);
        const int q0 = prog.e_first[e], q1 = prog.e_first[e + 1];
        u.assign(static_cast<size_t>(q1 - q0) * prog.max_nres, 0.0);
        for (int q = q0; q < q1; ++q) {
          const RBlock& rb = prog.rbs[prog.e_rbs[q]];
          double* uu = u.data() + static_cast<size_t>(q - q0) * prog.max_nres;
          for (size_t s = 0; s < rb.user.size(); ++s) {
            const int fid = rb.free_id[s];
            if (fid < 0 || prog.blocks[fid].is_e) continue;
            const FreeBlock& fb = prog.blocks[fid];
            const double* F = jac.data() + rb.jac_off[s];
            for (int k = 0; k < rb.nres; ++k)
              for (int c = 0; c < fb.size; ++c) uu[k] += F[k * fb.size + c] * p[fb.ef_offset + c];
          }
          const double* E = jac.data() + rb.jac_off[rb.e_slot];
          for (int k = 0; k < rb.nres; ++k)
            for (int i = 0; i < es; ++i) z[i] += E[k * es + i] * uu[k];
        }
        const double* inv = ete_inv.data() + ete_off[e];
        for (int i = 0; i < es; ++i)
          for (int j = 0; j < es; ++j) yv[i] += inv[i * es + j] * z[j];
        for (int q = q0; q < q1; ++q) {
          const RBlock& rb = prog.rbs[prog.e_rbs[q]];
          double* uu = u.data() + static_cast<size_t>(q - q0) * prog.max_nres;
          const double* E = jac.data() + rb.jac_off[rb.e_slot];
          for (int k = 0; k < rb.nres; ++k)
            for (int i = 0; i < es; ++i) uu[k] -= E[k * es + i] * yv[i];
          for (size_t s = 0; s < rb.user.size(); ++s) {
            const int fid = rb.free_id[s];
            if (fid < 0 || prog.blocks[fid].is_e) continue;
            const FreeBlock& fb = prog.blocks[fid];
            const double* F = jac.data() + rb.jac_off[s];
            for (int k = 0; k < rb.nres; ++k)
              for (int c = 0; c < fb.size; ++c) ql[fb.ef_offset + c] += F[k * fb.size + c] * uu[k];
          }
        }
      }
      const int nno = static_cast<int>(prog.no_e_rbs.size());
#pragma omp for schedule(static)
      for (int qi = 0; qi < nno; ++qi) {
        const RBlock& rb = prog.rbs[prog.no_e_rbs[qi]];
        double uu[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (size_t s = 0; s < rb.user.size(); ++s) {
          const int fid = rb.free_id[s];
          if (fid < 0) continue;
          const FreeBlock& fb = prog.blocks[fid];
          const double* F = jac.data() + rb.jac_off[s];
          for (int k = 0; k < rb.nres && k < 8; ++k)
            for (int c = 0; c < fb.size; ++c) uu[k] += F[k * fb.size + c] * p[fb.ef_offset + c];
        }
        for (size_t s = 0; s < rb.user.size(); ++s) {
          const int fid = rb.free_id[s];
          if (fid < 0) continue;
          const FreeBlock& fb = prog.blocks[fid];
          const double* F = jac.data() + rb.jac_off[s];
          for (int k = 0; k < rb.nres && k < 8; ++k)
            for (int c = 0; c < fb.size; ++c) ql[fb.ef_offset + c] += F[k * fb.size + c] * uu[k];
        }
      }
    }
    for (int i = 0; i < nf; ++i) {
      double acc = 0.0;
      for (int t = 0; t < nt; ++t) acc += qp[t][i];
      (*q)[i] = acc;
    }
    for (int f = 0; f < prog.num_f; ++f) {
      const FreeBlock& fb = prog.blocks[prog.f_block_ids[f]];
      for (int i = 0; i < fb.size; ++i)
        (*q)[fb.ef_offset + i] += D[fb.offset + i] * D[fb.offset + i] * p[fb.ef_offset + i];
    }
  };
  (void)nrb;
  std::vector<double> xs(nf, 0.0), r(b), z(nf, 0.0), p(nf, 0.0), q(nf, 0.0);
  apply_minv(r, &z);
  p = z;
  double rz = 0.0;
  for (int i = 0; i < nf; ++i) rz += r[i] * z[i];
  const double rz0 = rz;
  int it = 0;
  while (it < max_iter && nf > 0) {
    if (it >= min_iter && rz <= rel_tol * rel_tol * rz0) break;
    if (!(rz > 0.0)) break;
    apply_S(p, &q);
    double pq = 0.0;
    for (int i = 0; i < nf; ++i) pq += p[i] * q[i];
    if (!(pq > 0.0) || !std::isfinite(pq)) break;
    const double alpha = rz / pq;
    for (int i = 0; i < nf; ++i) {
      xs[i] += alpha * p[i];
      r[i] -= alpha * q[i];
    }
    apply_minv(r, &z);
    double rz_new = 0.0;
    for (int i = 0; i < nf; ++i) rz_new += r[i] * z[i];
    const double beta = rz_new / rz;
    rz = rz_new;
    for (int i = 0; i < nf; ++i) p[i] = z[i] + beta * p[i];
    ++it;
  }
  *iterations = it;
  for (int fbi : prog.f_block_ids) {
    const FreeBlock& fb = prog.blocks[fbi];
    for (int i = 0; i < fb.size; ++i) y[fb.offset + i] = xs[fb.ef_offset + i];
  }
  // back-substitute
#pragma omp parallel for schedule(dynamic, 64) num_threads(nt)
  for (int e = 0; e < ne_blocks; ++e) {
    const FreeBlock& eb = prog.blocks[prog.e_block_ids[e]];
    const int es = eb.size;
    std::vector<double> acc(es, 0.0);
    for (int qq = prog.e_first[e]; qq < prog.e_first[e + 1]; ++qq) {
      const RBlock& rb = prog.rbs[prog.e_rbs[qq]];
      const double* E = jac.data() + rb.jac_off[rb.e_slot];
      const double* res = residuals.data() + rb.row;
      for (int k = 0; k < rb.nres; ++k) {
        double t = res[k];
        for (size_t s = 0; s < rb.user.size(); ++s) {
          const int fid = rb.free_id[s];
          if (fid < 0 || prog.blocks[fid].is_e) continue;
          const FreeBlock& fb = prog.blocks[fid];
          const double* F = jac.data() + rb.jac_off[s];
          for (int c = 0; c < fb.size; ++c) t -= F[k * fb.size + c] * y[fb.offset + c];
        }
        for (int i = 0; i < es; ++i) acc[i] += E[k * es + i] * t;
      }
    }
    const double* inv = ete_inv.data() + ete_off[e];
    for (int i = 0; i < es; ++i) {
      double s = 0.0;
      for (int j = 0; j < es; ++j) s += inv[i * es + j] * acc[j];
      y[eb.offset + i] = s;
    }
  }
  return true;
}

Solver::Summary g_last_summary;
shim::Overrides g_overrides;

const char* LinearSolverName(LinearSolverType t) {
  switch (t) {
    case DENSE_SCHUR: return "DENSE_SCHUR";
    case ITERATIVE_SCHUR: return "ITERATIVE_SCHUR (shim: implicit Schur + block-Jacobi PCG)";
    case SPARSE_SCHUR: return "SPARSE_SCHUR";
    case DENSE_QR: return "DENSE_QR";
    case DENSE_NORMAL_CHOLESKY: return "DENSE_NORMAL_CHOLESKY";
    case SPARSE_NORMAL_CHOLESKY: return "SPARSE_NORMAL_CHOLESKY";
    case CGNR: return "CGNR";
  }
  return "?";
}

}  // namespace

namespace shim {
const Solver::Summary& LastSummary() { return g_last_summary; }
Overrides& GlobalOverrides() { return g_overrides; }
}  // namespace shim

// Trust-region LM [Ceres-upstream trust_region_minimizer.cc (2.x control flow),
// levenberg_marquardt_strategy.cc, trust_region_step_evaluator.cc (monotonic)].
void Solve(const Solver::Options& options_in, Problem* problem, Solver::Summary* summary) {
  Solver::Options options = options_in;
  const shim::Overrides& ov = g_overrides;
  if (ov.quiet == 1) options.minimizer_progress_to_stdout = false;
  if (ov.num_threads > 0) options.num_threads = ov.num_threads;
  if (ov.max_num_iterations >= 0) options.max_num_iterations = ov.max_num_iterations;
  if (ov.function_tolerance >= 0.0) options.function_tolerance = ov.function_tolerance;
  if (ov.gradient_tolerance >= 0.0) options.gradient_tolerance = ov.gradient_tolerance;
  if (ov.parameter_tolerance >= 0.0) options.parameter_tolerance = ov.parameter_tolerance;
  if (ov.linear_solver_type >= 0)
    options.linear_solver_type = static_cast<LinearSolverType>(ov.linear_solver_type);
  if (ov.max_linear_solver_iterations >= 0)
    options.max_linear_solver_iterations = ov.max_linear_solver_iterations;
  if (ov.min_linear_solver_iterations >= 0)
    options.min_linear_solver_iterations = ov.min_linear_solver_iterations;
  if (ov.shim_pcg_rel_tol >= 0.0) options.shim_pcg_rel_tol = ov.shim_pcg_rel_tol;

  const double t_start = Now();
  *summary = Solver::Summary();
  internal::ProblemImpl* pi = problem->impl();
  const int threads = std::max(1, std::min(options.num_threads, omp_get_num_procs()));
  summary->num_threads_given = options.num_threads;
  summary->num_threads_used = threads;
  summary->linear_solver_type_used = options.linear_solver_type;

  Program prog;
  // Every linear solver type yields the same exact step; the Schur split is used for
  // all of them except ITERATIVE_SCHUR, which runs the implicit PCG extension.
  Preprocess(pi, /*want_schur=*/true, &prog);
  summary->num_parameter_blocks = static_cast<int>(pi->params.size());
  for (const auto& p : pi->params) summary->num_parameters += p.size;
  summary->num_residual_blocks = static_cast<int>(pi->residuals.size());
  for (const auto& r : pi->residuals) summary->num_residuals += r.cost->num_residuals();
  summary->num_parameter_blocks_reduced = static_cast<int>(prog.blocks.size());
  summary->num_parameters_reduced = prog.num_params;
  summary->num_residual_blocks_reduced = static_cast<int>(prog.rbs.size());
  summary->num_residuals_reduced = prog.num_residuals;
  summary->num_e_blocks = prog.num_e;
  summary->num_f_blocks = prog.num_f;
  summary->reduced_system_size = prog.f_size;
  summary->fixed_cost = prog.fixed_cost;

  const int n = prog.num_params;
  std::vector<double> x(n), x_plus(n), gradient, residuals, cand_residuals, jac, scale(n, 1.0);
  for (const FreeBlock& b : prog.blocks)
    for (int i = 0; i < b.size; ++i) x[b.offset + i] = b.user[i];

  auto finish = [&](TerminationType tt, const std::string& msg, double cost) {
    summary->termination_type = tt;
    summary->message = msg;
    summary->final_cost = cost + prog.fixed_cost;
    summary->total_time_in_seconds = Now() - t_start;
    for (const FreeBlock& b : prog.blocks)
      for (int i = 0; i < b.size; ++i) b.user[i] = x[b.offset + i];
    g_last_summary = *summary;
  };

  if (n == 0 || prog.rbs.empty()) {
    summary->initial_cost = prog.fixed_cost;
    finish(CONVERGENCE, "Function tolerance reached. No non-constant parameter blocks found.", 0.0);
    return;
  }

  double x_cost = 0.0;
  double t0 = Now();
  bool jacobian_ok = Evaluate(prog, x.data(), true, &x_cost, &residuals, &jac, &gradient, threads);
  summary->jacobian_evaluation_time_in_seconds += Now() - t0;
  summary->num_jacobian_evaluations++;
  summary->num_residual_evaluations++;
  if (!jacobian_ok) {
    summary->initial_cost = x_cost + prog.fixed_cost;
    finish(FAILURE, "Initial residual and Jacobian evaluation failed.", x_cost);
    return;
  }
  if (options.jacobi_scaling) {
    // jacobian_scaling_[i] = 1 / (1 + sqrt(SquaredColumnNorm[i])), fixed at iteration 0
    SquaredColumnNorms(prog, jac, &scale);
    for (int i = 0; i < n; ++i) scale[i] = 1.0 / (1.0 + std::sqrt(scale[i]));
    ScaleColumns(prog, scale, &jac, threads);
  }
  auto gradient_norms = [&](double* max_norm, double* norm) {
    double m = 0.0, s = 0.0;
    for (int i = 0; i < n; ++i) {
      m = std::max(m, std::fabs(gradient[i]));
      s += gradient[i] * gradient[i];
    }
    *max_norm = m;
    *norm = std::sqrt(s);
  };
  auto vec_norm = [&](const std::vector<double>& v) {
    double s = 0.0;
    for (double a : v) s += a * a;
    return std::sqrt(s);
  };

  IterationSummary it;
  it.iteration = 0;
  it.step_is_valid = false;
  it.step_is_successful = false;
  it.cost = x_cost + prog.fixed_cost;
  gradient_norms(&it.gradient_max_norm, &it.gradient_norm);
  double x_norm = vec_norm(x);
  summary->initial_cost = it.cost;

  // LM strategy state
  double radius = options.initial_trust_region_radius;
  double decrease_factor = 2.0;
  bool reuse_diagonal = false;
  std::vector<double> diagonal(n), lm_diagonal(n), step(n), delta(n), model_residuals;
  int num_consecutive_invalid_steps = 0;
  SchurWork work;

  if (options.minimizer_progress_to_stdout) {
    std::printf(
        "iter      cost      cost_change  |gradient|   |step|    tr_ratio  tr_radius  ls_iter  "
        "iter_time  total_time\n");
  }
  double iter_start = t_start;
  bool first_finalize = true;
  // FinalizeIterationAndCheckIfMinimizerCanContinue
  auto finalize = [&]() -> bool {
    if (!first_finalize) {
      if (it.step_is_successful)
        summary->num_successful_steps++;
      else
        summary->num_unsuccessful_steps++;
    }
    first_finalize = false;
    it.trust_region_radius = radius;
    it.iteration_time_in_seconds = Now() - iter_start;
    it.cumulative_time_in_seconds = Now() - t_start;
    summary->iterations.push_back(it);
    if (options.minimizer_progress_to_stdout) {
      std::printf("%4d % 8e   % 3.2e   % 3.2e  % 3.2e  % 3.2e % 3.2e     % 4d   % 3.2e   % 3.2e\n",
                  it.iteration, it.cost, it.cost_change, it.gradient_max_norm, it.step_norm,
                  it.relative_decrease, it.trust_region_radius, it.linear_solver_iterations,
                  it.iteration_time_in_seconds, it.cumulative_time_in_seconds);
      std::fflush(stdout);
    }
    if (it.cumulative_time_in_seconds >= options.max_solver_time_in_seconds) {
      finish(NO_CONVERGENCE, "Maximum solver time reached.", x_cost);
      return false;
    }
    if (it.iteration >= options.max_num_iterations) {
      finish(NO_CONVERGENCE, "Maximum number of iterations reached.", x_cost);
      return false;
    }
    if (it.gradient_max_norm <= options.gradient_tolerance) {
      finish(CONVERGENCE, "Gradient tolerance reached.", x_cost);
      return false;
    }
    if (radius <= options.min_trust_region_radius) {
      finish(CONVERGENCE, "Minimum trust region radius reached.", x_cost);
      return false;
    }
    return true;
  };

  while (finalize()) {
    iter_start = Now();
    const IterationSummary prev = it;
    it = IterationSummary();
    it.iteration = prev.iteration + 1;
    it.eta = options.eta;

    // ---- ComputeTrustRegionStep (LevenbergMarquardtStrategy::ComputeStep)
    if (!reuse_diagonal) {
      SquaredColumnNorms(prog, jac, &diagonal);
      for (int i = 0; i < n; ++i)
        diagonal[i] = std::min(std::max(diagonal[i], options.min_lm_diagonal), options.max_lm_diagonal);
    }
    for (int i = 0; i < n; ++i) lm_diagonal[i] = std::sqrt(diagonal[i] / radius);
    t0 = Now();
    std::fill(step.begin(), step.end(), std::numeric_limits<double>::quiet_NaN());
    bool linear_ok;
    int ls_iters = 1;
    if (options.linear_solver_type == ITERATIVE_SCHUR) {
      linear_ok = SolveImplicit(prog, jac, residuals, lm_diagonal, threads,
                                options.min_linear_solver_iterations,
                                options.max_linear_solver_iterations, options.shim_pcg_rel_tol,
                                step.data(), &ls_iters);
    } else {
      linear_ok = EliminateAndSolve(prog, jac, residuals, lm_diagonal, threads, step.data(), &work);
    }
    summary->linear_solver_time_in_seconds += Now() - t0;
    summary->num_linear_solves++;
    it.linear_solver_iterations = ls_iters;
    if (linear_ok)
      for (int i = 0; i < n; ++i)
        if (!std::isfinite(step[i])) linear_ok = false;
    reuse_diagonal = true;
    bool step_valid = false;
    double model_cost_change = 0.0;
    if (linear_ok) {
      for (int i = 0; i < n; ++i) step[i] = -step[i];
      // model_cost_change = -(J step)^T (r + J step / 2)
      RightMultiply(prog, jac, step.data(), &model_residuals, threads);
      double acc = 0.0;
      for (int i = 0; i < prog.num_residuals; ++i)
        acc += model_residuals[i] * (residuals[i] + model_residuals[i] / 2.0);
      model_cost_change = -acc;
      step_valid = model_cost_change > 0.0;  // upstream: step_is_valid = (model_cost_change_ > 0.0)
    }
    it.model_cost_change = model_cost_change;
    it.step_is_valid = step_valid;
    if (!step_valid) {
      // HandleInvalidStep
      if (++num_consecutive_invalid_steps >= options.max_num_consecutive_invalid_steps) {
        std::ostringstream os;
        os << "Number of consecutive invalid steps more than "
              "Solver::Options::max_num_consecutive_invalid_steps: "
           << options.max_num_consecutive_invalid_steps;
        it.cost = x_cost + prog.fixed_cost;
        it.trust_region_radius = radius;
        summary->iterations.push_back(it);
        finish(FAILURE, os.str(), x_cost);
        return;
      }
      radius = radius / decrease_factor;  // StepIsInvalid() == StepRejected(0)
      decrease_factor *= 2.0;
      reuse_diagonal = true;
      it.cost = x_cost + prog.fixed_cost;
      it.cost_change = 0.0;
      it.gradient_max_norm = prev.gradient_max_norm;
      it.gradient_norm = prev.gradient_norm;
      it.step_norm = 0.0;
      it.relative_decrease = 0.0;
      continue;
    }
    num_consecutive_invalid_steps = 0;

    // ---- ComputeCandidatePointAndEvaluateCost
    for (int i = 0; i < n; ++i) {
      delta[i] = step[i] * scale[i];
      x_plus[i] = x[i] + delta[i];
    }
    double candidate_cost;
    t0 = Now();
    const bool eval_ok =
        Evaluate(prog, x_plus.data(), false, &candidate_cost, &cand_residuals, NULL, NULL, threads);
    summary->residual_evaluation_time_in_seconds += Now() - t0;
    summary->num_residual_evaluations++;
    if (!eval_ok) candidate_cost = std::numeric_limits<double>::max();

    // ---- ParameterToleranceReached
    double sn = 0.0;
    for (int i = 0; i < n; ++i) sn += (x[i] - x_plus[i]) * (x[i] - x_plus[i]);
    it.step_norm = std::sqrt(sn);
    const double step_size_tolerance = options.parameter_tolerance * (x_norm + options.parameter_tolerance);
    if (it.step_norm <= step_size_tolerance) {
      std::ostringstream os;
      os << "Parameter tolerance reached. Relative step_norm: "
         << it.step_norm / (x_norm + options.parameter_tolerance) << " <= " << options.parameter_tolerance;
      it.cost = x_cost + prog.fixed_cost;
      it.trust_region_radius = radius;
      summary->iterations.push_back(it);
      finish(CONVERGENCE, os.str(), x_cost);
      return;
    }
    // ---- FunctionToleranceReached
    it.cost_change = x_cost - candidate_cost;
    const double absolute_function_tolerance = options.function_tolerance * x_cost;
    if (std::fabs(it.cost_change) <= absolute_function_tolerance) {
      std::ostringstream os;
      os << "Function tolerance reached. |cost_change|/cost: " << std::fabs(it.cost_change) / x_cost
         << " <= " << options.function_tolerance;
      it.cost = x_cost + prog.fixed_cost;
      it.trust_region_radius = radius;
      summary->iterations.push_back(it);
      finish(CONVERGENCE, os.str(), x_cost);
      return;
    }
    // ---- IsStepSuccessful (monotonic TrustRegionStepEvaluator::StepQuality)
    if (candidate_cost >= std::numeric_limits<double>::max())
      it.relative_decrease = std::numeric_limits<double>::lowest();
    else
      it.relative_decrease = (x_cost - candidate_cost) / model_cost_change;
    if (it.relative_decrease > options.min_relative_decrease) {
      // HandleSuccessfulStep
      x = x_plus;
      x_norm = vec_norm(x);
      t0 = Now();
      const bool ok2 = Evaluate(prog, x.data(), true, &x_cost, &residuals, &jac, &gradient, threads);
      summary->jacobian_evaluation_time_in_seconds += Now() - t0;
      summary->num_jacobian_evaluations++;
      summary->num_residual_evaluations++;
      if (!ok2) {
        finish(FAILURE, "Residual and Jacobian evaluation failed.", x_cost);
        return;
      }
      if (options.jacobi_scaling) ScaleColumns(prog, scale, &jac, threads);
      it.cost = x_cost + prog.fixed_cost;
      gradient_norms(&it.gradient_max_norm, &it.gradient_norm);
      it.step_is_successful = true;
      // StepAccepted
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * it.relative_decrease - 1.0, 3));
      radius = std::min(options.max_trust_region_radius, radius);
      decrease_factor = 2.0;
      reuse_diagonal = false;
    } else {
      it.step_is_successful = false;
      it.cost = candidate_cost + prog.fixed_cost;
      it.gradient_max_norm = prev.gradient_max_norm;
      it.gradient_norm = prev.gradient_norm;
      // StepRejected
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      reuse_diagonal = true;
    }
  }
}

std::string Solver::Summary::BriefReport() const {
  std::ostringstream os;
  os << "mini-Ceres Solver Report: Iterations: " << iterations.size()
     << ", Initial cost: " << initial_cost << ", Final cost: " << final_cost << ", Termination: "
     << (termination_type == CONVERGENCE ? "CONVERGENCE"
                                         : termination_type == NO_CONVERGENCE ? "NO_CONVERGENCE" : "FAILURE");
  return os.str();
}

std::string Solver::Summary::FullReport() const {
  std::ostringstream os;
  char buf[256];
  os << "\nSolver Summary (mini-Ceres oracle restatement; NOT ceres-solver)\n\n";
  os << "                                     Original                  Reduced\n";
  std::snprintf(buf, sizeof buf, "Parameter blocks    %25d%25d\n", num_parameter_blocks, num_parameter_blocks_reduced);
  os << buf;
  std::snprintf(buf, sizeof buf, "Parameters          %25d%25d\n", num_parameters, num_parameters_reduced);
  os << buf;
  std::snprintf(buf, sizeof buf, "Residual blocks     %25d%25d\n", num_residual_blocks, num_residual_blocks_reduced);
  os << buf;
  std::snprintf(buf, sizeof buf, "Residuals           %25d%25d\n\n", num_residuals, num_residuals_reduced);
  os << buf;
  os << "Minimizer                        TRUST_REGION\n";
  os << "Trust region strategy     LEVENBERG_MARQUARDT\n\n";
  os << "Linear solver          " << LinearSolverName(linear_solver_type_used) << "\n";
  os << "Threads (given/used)   " << num_threads_given << " / " << num_threads_used << "\n";
  os << "Schur structure        e-blocks " << num_e_blocks << ", f-blocks " << num_f_blocks
     << ", reduced system " << reduced_system_size << "\n\n";
  os << "Cost:\n";
  std::snprintf(buf, sizeof buf, "Initial          %30e\nFinal            %30e\nChange           %30e\n\n",
                initial_cost, final_cost, initial_cost - final_cost);
  os << buf;
  os << "Minimizer iterations " << iterations.size() << "\n";
  os << "Successful steps     " << num_successful_steps << "\n";
  os << "Unsuccessful steps   " << num_unsuccessful_steps << "\n\n";
  os << "Time (in seconds):\n";
  std::snprintf(buf, sizeof buf,
                "  Residual only evaluation %14.6f (%d)\n  Jacobian & residual evaluation %8.6f (%d)\n"
                "  Linear solver     %21.6f (%d)\nTotal              %22.6f\n\n",
                residual_evaluation_time_in_seconds, num_residual_evaluations,
                jacobian_evaluation_time_in_seconds, num_jacobian_evaluations,
                linear_solver_time_in_seconds, num_linear_solves, total_time_in_seconds);
  os << buf;
  os << "Termination: "
     << (termination_type == CONVERGENCE ? "CONVERGENCE"
                                         : termination_type == NO_CONVERGENCE ? "NO_CONVERGENCE" : "FAILURE")
     << " (" << message << ")\n";
  return os.str();
}

}  // namespace ceres
