// Hemisphere-radius fit on the device: the whole Levenberg-Marquardt loop of the reference's
// second Ceres problem (reference src/sfm.cc:86-103) runs inside ONE kernel launch.
//
//   residual_k = |c - pos_k|^2 - rho          (src/hemisphere_radius.hh:18-28; rho = radius^2)
//   unknowns   c (3), rho (1); init 0,0,0 / 1  (src/sfm.cc:87-88)
//   jacobian   d r_k / d c = 2 (c - pos_k),  d r_k / d rho = -1
// The problem has 4 unknowns and a few hundred residuals (one per rig camera), so it is
// latency-, not bandwidth-bound: one CTA keeps the normal equations in registers/shared
// memory, block-reduces J^T J, J^T r and the cost each iteration, and thread 0 runs the
// trust-region logic (same restatement of Ceres' LM as ba_engine.cu / oracle/mini_ceres.cc:
// Jacobi scaling fixed at iteration 0, D^2 = clamp(diag J^T J)/radius, exact solve,
// rho = cost change / model change, radius schedule, Ceres 2.x termination order).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/deeparc_ba.h"

namespace {

struct HemiOptions {
  int max_iter, max_invalid, jacobi_scaling;
  double max_seconds;
  double radius0, max_radius, min_radius, min_rel_decrease, min_diag, max_diag, ftol, gtol, ptol;
};

struct HemiResult {
  double x[4];
  double initial_cost, final_cost;
  int termination, n_iterations, n_success, n_fail, reason;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Reduces NV values per thread across the CTA; every thread gets the totals in out[].
template <int NV>
__device__ void block_reduce(double (&v)[NV], double* smem /* [8][NV] */, double (&out)[NV]) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double s = warp_sum(v[k]);
    if (lane == 0) smem[wid * NV + k] = s;
  }
  __syncthreads();
  const int nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double s = 0.0;
    for (int w = 0; w < nw; ++w) s += smem[w * NV + k];
    out[k] = s;
  }
}

// Solves the SPD 4x4 system A y = b by Cholesky; returns false if A is not positive definite.
__device__ bool solve4(const double A[4][4], const double b[4], double y[4]) {
  double L[4][4];
  for (int j = 0; j < 4; ++j) {
    double d = A[j][j];
    for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
    if (!(d > 0.0)) return false;
    d = sqrt(d);
    L[j][j] = d;
    for (int i = j + 1; i < 4; ++i) {
      double s = A[i][j];
      for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
      L[i][j] = s / d;
    }
  }
  double z[4];
  for (int i = 0; i < 4; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= L[i][k] * z[k];
    z[i] = s / L[i][i];
  }
  for (int i = 3; i >= 0; --i) {
    double s = z[i];
    for (int k = i + 1; k < 4; ++k) s -= L[k][i] * y[k];
    y[i] = s / L[i][i];
  }
  return true;
}

// Every thread runs the (scalar, deterministic) LM logic redundantly on the block-reduced
// totals, so no control state has to be broadcast; thread 0 alone writes the records.
__global__ void __launch_bounds__(256) k_hemisphere_lm(const double* __restrict__ pos, int n, HemiOptions o,
                                                        double x0, double x1, double x2, double x3,
                                                        HemiResult* __restrict__ result,
                                                        dba_iteration* __restrict__ iters, int iters_cap) {
  __shared__ double red[8 * 15];
  __shared__ int sh_timeout;
  double x[4] = {x0, x1, x2, x3};
  double scale[4] = {1.0, 1.0, 1.0, 1.0};
  double radius = o.radius0, decrease_factor = 2.0, x_cost = 0.0, x_norm = 0.0, gmax = 0.0, gnorm = 0.0;
  double H[4][4], g[4];
  int n_it = 0, n_success = 0, n_fail = 0, invalid = 0, termination = DBA_NO_CONVERGENCE, reason = 0;
  unsigned long long t0 = 0;
  if (threadIdx.x == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));

  auto record = [&](const dba_iteration& it) {
    if (threadIdx.x == 0 && n_it < iters_cap) iters[n_it] = it;
    ++n_it;
  };
  // normal equations of the scaled Jacobian at x (10 unique H entries, 4 g entries, sum r^2)
  auto linearise = [&](bool fix_scaling) {
    double v[15], colsq[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 15; ++k) v[k] = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const double d0 = x[0] - pos[3 * i], d1 = x[1] - pos[3 * i + 1], d2 = x[2] - pos[3 * i + 2];
      const double r = d0 * d0 + d1 * d1 + d2 * d2 - x[3];
      double J[4] = {2.0 * d0, 2.0 * d1, 2.0 * d2, -1.0};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        colsq[k] += J[k] * J[k];
        J[k] *= scale[k];
      }
      int u = 0;
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a; b < 4; ++b) v[u++] += J[a] * J[b];
#pragma unroll
      for (int a = 0; a < 4; ++a) v[10 + a] += J[a] * r;
      v[14] += r * r;
    }
    if (fix_scaling) {
      // jacobian_scaling = 1 / (1 + ||column||), fixed at iteration 0; scale == 1 so far
      double cs[4];
      block_reduce<4>(colsq, red, cs);
#pragma unroll
      for (int k = 0; k < 4; ++k) scale[k] = 1.0 / (1.0 + sqrt(cs[k]));
      int u = 0;
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a; b < 4; ++b) v[u++] *= scale[a] * scale[b];
#pragma unroll
      for (int a = 0; a < 4; ++a) v[10 + a] *= scale[a];
    }
    double tot[15];
    block_reduce<15>(v, red, tot);
    int u = 0;
    for (int a = 0; a < 4; ++a)
      for (int b = a; b < 4; ++b) {
        H[a][b] = tot[u];
        H[b][a] = tot[u];
        ++u;
      }
    gmax = 0.0;
    gnorm = 0.0;
    for (int a = 0; a < 4; ++a) {
      g[a] = tot[10 + a];
      const double gu = g[a] / scale[a];
      gmax = fmax(gmax, fabs(gu));
      gnorm += gu * gu;
    }
    gnorm = sqrt(gnorm);
    x_cost = 0.5 * tot[14];
    x_norm = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
  };

  linearise(o.jacobi_scaling != 0);
  const double initial_cost = x_cost;
  dba_iteration it;
  memset(&it, 0, sizeof it);
  it.cost = x_cost;
  it.gradient_max_norm = gmax;
  it.gradient_norm = gnorm;
  bool first = true;
  for (;;) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue
    if (!first) {
      if (it.step_is_successful) ++n_success; else ++n_fail;
    }
    first = false;
    it.trust_region_radius = radius;
    record(it);
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
      sh_timeout = ((t1 - t0) * 1e-9 >= o.max_seconds) ? 1 : 0;
    }
    __syncthreads();
    if (sh_timeout) { termination = DBA_NO_CONVERGENCE; reason = 1; break; }
    if (it.iteration >= o.max_iter) { termination = DBA_NO_CONVERGENCE; reason = 2; break; }
    if (it.gradient_max_norm <= o.gtol) { termination = DBA_CONVERGENCE; reason = 3; break; }
    if (radius <= o.min_radius) { termination = DBA_CONVERGENCE; reason = 4; break; }

    const dba_iteration prev = it;
    memset(&it, 0, sizeof it);
    it.iteration = prev.iteration + 1;
    it.linear_solver_iterations = 1;
    // (H + D^2) y = g, step = -y;  model change = -(step.g + step^T H step / 2)
    double A[4][4], y[4], step[4];
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b) A[a][b] = H[a][b];
    for (int a = 0; a < 4; ++a) A[a][a] += fmin(fmax(H[a][a], o.min_diag), o.max_diag) / radius;
    bool ok = solve4(A, g, y);
    double model_change = 0.0;
    if (ok) {
      double q = 0.0, lin = 0.0;
      for (int a = 0; a < 4; ++a) {
        step[a] = -y[a];
        lin += step[a] * g[a];
      }
      for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) q += step[a] * H[a][b] * step[b];
      model_change = -(lin + 0.5 * q);
      ok = isfinite(model_change) && model_change > 0.0;
    }
    it.model_cost_change = model_change;
    it.step_is_valid = ok ? 1 : 0;
    if (!ok) {  // HandleInvalidStep
      it.cost = x_cost;
      it.gradient_max_norm = prev.gradient_max_norm;
      it.gradient_norm = prev.gradient_norm;
      if (++invalid >= o.max_invalid) {
        it.trust_region_radius = radius;
        record(it);
        termination = DBA_FAILURE;
        reason = 5;
        break;
      }
      radius /= decrease_factor;
      decrease_factor *= 2.0;
      continue;
    }
    invalid = 0;
    double xc[4], sn = 0.0;
    for (int a = 0; a < 4; ++a) {
      const double d = scale[a] * step[a];
      xc[a] = x[a] + d;
      sn += d * d;
    }
    double cv[1] = {0.0}, ct[1];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const double d0 = xc[0] - pos[3 * i], d1 = xc[1] - pos[3 * i + 1], d2 = xc[2] - pos[3 * i + 2];
      const double r = d0 * d0 + d1 * d1 + d2 * d2 - xc[3];
      cv[0] += r * r;
    }
    block_reduce<1>(cv, red, ct);
    double cand = 0.5 * ct[0];
    if (!isfinite(cand)) cand = 1.7976931348623157e308;
    it.step_norm = sqrt(sn);
    if (it.step_norm <= o.ptol * (x_norm + o.ptol)) {  // ParameterToleranceReached
      it.cost = x_cost;
      it.trust_region_radius = radius;
      record(it);
      termination = DBA_CONVERGENCE;
      reason = 6;
      break;
    }
    it.cost_change = x_cost - cand;
    if (fabs(it.cost_change) <= o.ftol * x_cost) {  // FunctionToleranceReached
      it.cost = x_cost;
      it.trust_region_radius = radius;
      record(it);
      termination = DBA_CONVERGENCE;
      reason = 7;
      break;
    }
    it.relative_decrease = cand >= 1.7976931348623157e308 ? -1.7976931348623157e308 : (x_cost - cand) / model_change;
    if (it.relative_decrease > o.min_rel_decrease) {
      for (int a = 0; a < 4; ++a) x[a] = xc[a];
      linearise(false);
      it.cost = x_cost;
      it.gradient_max_norm = gmax;
      it.gradient_norm = gnorm;
      it.step_is_successful = 1;
      const double t = 2.0 * it.relative_decrease - 1.0;
      radius = fmin(o.max_radius, radius / fmax(1.0 / 3.0, 1.0 - t * t * t));
      decrease_factor = 2.0;
    } else {
      it.cost = cand;
      it.gradient_max_norm = prev.gradient_max_norm;
      it.gradient_norm = prev.gradient_norm;
      radius /= decrease_factor;
      decrease_factor *= 2.0;
    }
  }
  if (threadIdx.x == 0) {
    for (int a = 0; a < 4; ++a) result->x[a] = x[a];
    result->initial_cost = initial_cost;
    result->final_cost = x_cost;
    result->termination = termination;
    result->n_iterations = n_it;
    result->n_success = n_success;
    result->n_fail = n_fail;
    result->reason = reason;
  }
}

const char* reason_text(int r) {
  switch (r) {
    case 1: return "Maximum solver time reached.";
    case 2: return "Maximum number of iterations reached.";
    case 3: return "Gradient tolerance reached.";
    case 4: return "Minimum trust region radius reached.";
    case 5: return "Number of consecutive invalid steps more than Solver::Options::max_num_consecutive_invalid_steps.";
    case 6: return "Parameter tolerance reached.";
    case 7: return "Function tolerance reached.";
  }
  return "";
}

}  // namespace

int hemisphere_fit_device(const double* centres_host, int n, double centre_io[3], double* rho_io,
                          const dba_solve_options* o, dba_summary* s, cudaStream_t st, std::string* err,
                          int64_t* launches) {
  double* d_pos = nullptr;
  HemiResult* d_res = nullptr;
  dba_iteration* d_it = nullptr;
  auto fail = [&](cudaError_t e, const char* what) {  // every error exit releases what was allocated so far
    *err = std::string(what) + ": " + cudaGetErrorString(e);
    cudaFree(d_pos);
    cudaFree(d_res);
    cudaFree(d_it);
    d_pos = nullptr;
    d_res = nullptr;
    d_it = nullptr;
    return DBA_ERR_CUDA;
  };
  dba_iteration* it_buf = s->iterations;
  const int cap = it_buf ? s->iterations_capacity : 0;
  std::memset(s, 0, sizeof *s);
  s->iterations = it_buf;
  s->iterations_capacity = cap;
  s->linear_solver_used = DBA_LS_DENSE;
  s->reduced_system_size = 4;
  s->termination = DBA_FAILURE;  // until the kernel has reported
  cudaError_t e;
  const int dev_cap = std::max(cap, 1);
  if ((e = cudaMalloc(&d_pos, sizeof(double) * 3 * std::max(n, 1))) != cudaSuccess) return fail(e, "cudaMalloc");
  if ((e = cudaMalloc(&d_res, sizeof(HemiResult))) != cudaSuccess) return fail(e, "cudaMalloc");
  if ((e = cudaMalloc(&d_it, sizeof(dba_iteration) * dev_cap)) != cudaSuccess) return fail(e, "cudaMalloc");
  if (n > 0) cudaMemcpyAsync(d_pos, centres_host, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, st);
  HemiOptions ho;
  ho.max_iter = o->max_num_iterations;
  ho.max_invalid = o->max_num_consecutive_invalid_steps;
  ho.jacobi_scaling = o->jacobi_scaling;
  ho.max_seconds = o->max_solver_time_in_seconds;
  ho.radius0 = o->initial_trust_region_radius;
  ho.max_radius = o->max_trust_region_radius;
  ho.min_radius = o->min_trust_region_radius;
  ho.min_rel_decrease = o->min_relative_decrease;
  ho.min_diag = o->min_lm_diagonal;
  ho.max_diag = o->max_lm_diagonal;
  ho.ftol = o->function_tolerance;
  ho.gtol = o->gradient_tolerance;
  ho.ptol = o->parameter_tolerance;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a, st);
  k_hemisphere_lm<<<1, 256, 0, st>>>(d_pos, n, ho, centre_io[0], centre_io[1], centre_io[2], *rho_io, d_res, d_it, cap);
  *launches = 1;
  cudaEventRecord(b, st);
  HemiResult res;
  std::vector<dba_iteration> its(dev_cap);
  cudaMemcpyAsync(&res, d_res, sizeof res, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(its.data(), d_it, sizeof(dba_iteration) * dev_cap, cudaMemcpyDeviceToHost, st);
  e = cudaStreamSynchronize(st);
  float ms = 0.f;
  if (e == cudaSuccess) cudaEventElapsedTime(&ms, a, b);
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(d_pos);
  cudaFree(d_res);
  cudaFree(d_it);
  if (e != cudaSuccess) return fail(e, "hemisphere kernel");
  centre_io[0] = res.x[0];
  centre_io[1] = res.x[1];
  centre_io[2] = res.x[2];
  *rho_io = res.x[3];
  s->termination = res.termination;
  s->num_iterations = std::min(res.n_iterations, cap);
  s->num_successful_steps = res.n_success;
  s->num_unsuccessful_steps = res.n_fail;
  s->initial_cost = res.initial_cost;
  s->final_cost = res.final_cost;
  s->device_time_in_seconds = ms * 1e-3;
  s->total_time_in_seconds = ms * 1e-3;
  s->kernel_launches = 1;
  std::snprintf(s->message, sizeof s->message, "%s", reason_text(res.reason));
  for (int i = 0; i < s->num_iterations; ++i) it_buf[i] = its[i];
  return DBA_OK;
}
