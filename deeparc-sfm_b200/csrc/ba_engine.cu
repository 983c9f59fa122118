// Host side of the engine behind the C ABI (include/deeparc_ba.h): problem upload and
// re-ordering, the GPU-resident Levenberg-Marquardt driver, parameter read-back.
//
// The LM control flow restates ceres::Solve as the reference configures it
// (reference src/sfm.cc:66-73; upstream ceres-solver 2.x trust_region_minimizer.cc and
// levenberg_marquardt_strategy.cc, see oracle/mini_ceres.cc for the CPU restatement the
// parity tests compare against).  All O(n_obs)/O(n_pts) work runs in the kernels of
// ba_kernels.cu; the host only sequences launches and reads ~20 scalars per LM iteration.
// There is no CPU fallback: any CUDA failure is reported through the status code.
#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "../../include/deeparc_ba.h"
#include "ba_build.cuh"
#include "ba_kernels.cuh"
#include "nccl_dyn.h"

using namespace dba;

int hemisphere_fit_device(const double* centres_host, int n, double centre_io[3], double* rho_io,
                          const dba_solve_options* o, dba_summary* s, cudaStream_t st, std::string* err,
                          int64_t* launches);

namespace {

thread_local std::string g_create_error;

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// scalar buffer layout (see DESIGN.md "LM scalars")
enum {
  S_COST = 0, S_CAND = 1, S_GSQ_PT = 2, S_MODEL = 3, S_STEP_PT = 4, S_X_PT = 5, S_BAD_PT = 6,  // sum over ranks
  S_GMAX_PT = 8,                                                                                 // max over ranks
  S_GSQ_CAM = 12, S_GMAX_CAM = 13, S_BAD_CAM = 14, S_STEP_CAM = 15, S_X_CAM = 16,                // replicated
  S_TOTAL = 24
};

struct KernelRecord {
  std::string name;
  int64_t launches = 0;
  double total_ms = 0.0;
  double algorithmic_bytes = 0.0;
};

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaError_t alloc(size_t count) {
    free();
    n = count;
    if (count == 0) return cudaSuccess;
    return cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T));
  }
  void free() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { free(); }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};

}  // namespace

struct dba_upload_state;
struct dba_handle {
  int device = 0, rank = 0, world = 1, verbose = 0;
  cudaStream_t st = nullptr;
  NcclComm comm = nullptr;
  std::string error;
  bool have_problem = false;

  // problem (host copies needed later)
  int64_t n_obs_global = 0, n_obs = 0;
  int n_pts_global = 0, n_pts = 0, pt_lo = 0, n_ext = 0, n_intr = 0;
  int cb = 0, two = 0, freeze = 0, free_intr = 0;
  std::vector<int64_t> perm;  // local sorted position -> caller's observation index
  bool any_const = false;

  DeviceProblem D{};
  ParamSet P[2]{};
  int cur = 0;
  WorkArrays W{};

  // device storage
  DevBuf<double2> d_obs_xy, d_J;
  DevBuf<int> d_cam_part_first;
  DevBuf<unsigned short> d_items, d_obs_lp, d_part_first_rel;
  DevBuf<TileMeta> d_tile_meta;
  DevBuf<int2> d_obs_ab;
  DevBuf<double> d_partials_q, d_q_split;
  int q_split = 1;  // slices of the per-camera partial sum (few camera blocks => long lists)
  // matrix-free implicit Schur product (default); DBA_SPMV=planes selects the product that reads
  // the materialised Jacobian planes (kept for A/B measurements)
  int mf = 1;
  int fuse_pcg = 1;  // one cooperative launch for the PCG vector work (DBA_PCG_FUSED=0 disables)
  // multi-GPU: peer windows of the fused PCG tail (CUDA IPC over NVLink); DBA_P2P=0 keeps the
  // ncclAllReduce path (kept for A/B measurements and for boxes without peer access)
  // the planes / cost partials / camera rows currently describe the CANDIDATE (speculative evaluation)
  bool pcg_pending = false;  // h_pcg_state is in flight (valid after the next stream synchronisation)
  bool jacobian_at_candidate = false;
  bool planeless = false;  // this solve keeps no camera planes: gather, product and back-substitution recompute F (set per dba_solve)
  int mf_front = 1;   // DBA_MF_FRONT=0: the camera-side gather reads the Jacobian planes instead of recomputing (single-pose problems)
  int mf_tail = 1;    // DBA_MF_TAIL=0: PCG tail as a separate launch (k_pcg_fused) instead of the k_spmv_mf epilogue
  int speculate = 1;  // DBA_SPECULATE=0: always evaluate candidates with the residual-only kernel
  bool p2p_ready = false;
  void* win_local = nullptr;
  size_t win_bytes = 0, win_slot_cap = 0;
  void* win_peer[kMaxPeers] = {};
  PeerWin pw{};
  unsigned long long p2p_seq = 0;
  DevBuf<unsigned char> d_ipc;
  DevBuf<int> d_mf_cols, d_part_dst, d_part_blk, d_items_mf, d_part_first;
  DevBuf<ushort2> d_obs_lc;
  DevBuf<double> d_mf_rows, d_mf_T;
  DevBuf<int2> d_obs_ip;
  DevBuf<int> d_tile_obs, d_tile_pt, d_pt_first, d_cam_entries, d_cam_chunk_first, d_nf, d_nd, d_pcg_state;
  DevBuf<unsigned int> d_counters;
  DevBuf<unsigned long long> d_trace;  // DBA_TAIL_TRACE=1
  // device-side problem construction (ba_build.cu): raw shard arrays and scratch
  DevBuf<double> d_raw_xy;
  DevBuf<int> d_raw_pt, d_raw_a, d_raw_b, d_raw_in, d_bld_a, d_bld_flags, d_bld_np, d_bld_g0, d_bld_key, d_bld_id, d_bld_key2, d_bld_id2,
      d_bld_first, d_bld_cnt;
  DevBuf<unsigned char> d_bld_temp;
  std::vector<double2> keep_xy_v;
  std::vector<int2> keep_ip_v, keep_ab_v;
  int build_mode = 0;  // how the last dba_problem_set built its index structures: 0 host cores, 1 device
  DevBuf<int4> d_cam_chunks;
  DevBuf<uint8_t> d_ext_const;
  DevBuf<double> d_center, d_pts[3], d_rot[3], d_trans[3], d_focal[3], d_dist[3];  // [2] = initial copy
  DevBuf<PoseRow> d_pose_rows[2];
  DevBuf<IntrRow> d_intr_rows[2];
  DevBuf<double> d_sp, d_sc, d_cinv, d_tp, d_dp, d_cam_acc, d_minv, d_dc2, d_x, d_r, d_z, d_p, d_q;
  DevBuf<double> d_cam_chunk_acc;
  DevBuf<double> d_partA, d_partB, d_scalars, d_scalars_red, d_pcg_scal, d_full_pts, d_vec_partials;
  // retained for dba_problem_update: the point-sorted observation arrays of the last dba_problem_set
  // (they live in the pinned staging arena until the next upload) and the small constant tables
  struct Keep {
    bool valid = false;
    const double2* xy = nullptr;
    const int2* ip = nullptr;  // (intrinsic, local point)
    const int2* ab = nullptr;  // (block a, block b or -1)
    std::vector<double> center;
    std::vector<int32_t> nf, nd;
    std::vector<uint8_t> ext_const;
    int free_intrinsics = 0;
  } keep;
  // explicit reduced system + device Cholesky (DENSE_SCHUR) for small camera counts
  bool dense_ok = false;     // the problem fits the dense path (set by dba_problem_set)
  bool use_dense = false;    // linear solver of the running dba_solve
  DenseWork Q{};
  DevBuf<int> d_dn_batch, d_dn_pair_entries, d_dn_pair_chunk_first;
  DevBuf<int4> d_dn_pair_chunks;
  DevBuf<double> d_dn_S, d_dn_Spart, d_dn_pair_acc, d_dn_Z;
  DevBuf<unsigned int> d_dn_Zent;
  DevBuf<int> d_dn_Zcount;
  int dense_failures = 0, pcg_unconverged = 0;  // per dba_solve
  double* h_scalars = nullptr;  // pinned
  int* h_pcg_state = nullptr;   // pinned
  struct dba_upload_state* upload = nullptr;  // pinned staging arena (grow-only)
  int64_t n_cam_entries = 0;
  size_t j_planes = 0;
  int plane_w = 0;  // extra plane holding w of the two-phase Schur product

  // accounting
  int64_t launches = 0;
  bool stats_enabled = false;
  std::vector<KernelRecord> records;
  std::map<std::string, int> record_index;
  struct Pending {
    int rec;
    cudaEvent_t a, b;
  };
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> event_pool;
  cudaEvent_t ev_solve[3] = {nullptr, nullptr, nullptr};  // start / end / loop start of dba_solve (created once)

  int fail(int status, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    error = buf;
    return status;
  }
};

#define CU(h, call)                                                                                   \
  do {                                                                                                \
    cudaError_t e__ = (call);                                                                         \
    if (e__ != cudaSuccess)                                                                           \
      return (h)->fail(DBA_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

// ---- upload helpers: grow-only pinned staging arena, grow-only device buffers
struct PinnedArena {
  char* base = nullptr;
  size_t cap = 0, used = 0;
  cudaError_t reserve(size_t bytes) {
    used = 0;
    if (bytes <= cap) return cudaSuccess;
    if (base) cudaFreeHost(base);
    base = nullptr;
    cap = 0;
    cudaError_t e = cudaMallocHost(reinterpret_cast<void**>(&base), bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  template <typename T>
  T* take(size_t n) {
    used = (used + 255) & ~size_t{255};
    T* p = reinterpret_cast<T*>(base + used);
    used += n * sizeof(T);
    return p;
  }
  ~PinnedArena() {
    if (base) cudaFreeHost(base);
  }
};
struct dba_upload_state {
  PinnedArena arena;     // staging of dba_problem_set; its observation arrays stay valid for dba_problem_update
  PinnedArena readback;  // staging of dba_params_get (separate: it must not overwrite the retained image)
};

namespace {
PinnedArena& arena_of(dba_handle* h) {
  if (!h->upload) h->upload = new dba_upload_state;
  return h->upload->arena;
}
template <typename T>
cudaError_t ensure(DevBuf<T>& b, size_t count) {
  if (count < 1) count = 1;
  if (b.p && b.n >= count) return cudaSuccess;
  return b.alloc(count);
}
inline size_t pad256(size_t bytes) { return (bytes + 255) & ~size_t{255}; }
}  // namespace

namespace {

// ---- kernel accounting ------------------------------------------------------------
int record_id(dba_handle* h, const char* name, double bytes) {
  auto it = h->record_index.find(name);
  if (it != h->record_index.end()) {
    if (bytes > 0) h->records[it->second].algorithmic_bytes = bytes;
    return it->second;
  }
  KernelRecord r;
  r.name = name;
  r.algorithmic_bytes = bytes;
  h->records.push_back(r);
  h->record_index[name] = static_cast<int>(h->records.size()) - 1;
  return static_cast<int>(h->records.size()) - 1;
}

cudaEvent_t get_event(dba_handle* h) {
  if (!h->event_pool.empty()) {
    cudaEvent_t e = h->event_pool.back();
    h->event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

struct Scope {
  dba_handle* h;
  int rec;
  cudaEvent_t a = nullptr, b = nullptr;
  Scope(dba_handle* hh, const char* name, double bytes = 0.0, int n_launch = 1) : h(hh) {
    rec = record_id(h, name, bytes);
    h->records[rec].launches += n_launch;
    h->launches += n_launch;
    if (h->stats_enabled) {
      a = get_event(h);
      b = get_event(h);
      cudaEventRecord(a, h->st);
    }
  }
  ~Scope() {
    if (a) {
      cudaEventRecord(b, h->st);
      h->pending.push_back({rec, a, b});
    }
  }
};

void drain_events(dba_handle* h) {
  for (auto& p : h->pending) {
    cudaEventSynchronize(p.b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, p.a, p.b);
    h->records[p.rec].total_ms += ms;
    h->event_pool.push_back(p.a);
    h->event_pool.push_back(p.b);
  }
  h->pending.clear();
}

// ---- collectives --------------------------------------------------------------------
int allreduce(dba_handle* h, double* buf, size_t n, int op) {
  if (h->world == 1 || n == 0) return DBA_OK;
  int rc = nccl_api().AllReduce(buf, buf, n, kNcclFloat64, op, h->comm, h->st);
  if (rc != 0) return h->fail(DBA_ERR_NCCL, "ncclAllReduce: %s", nccl_api().GetErrorString(rc));
  return DBA_OK;
}

// The local scalars stay untouched (several fetches may happen between two writes of a slot);
// the cross-rank combination goes to a second buffer that is what the host reads.
int reduce_scalars_and_fetch(dba_handle* h) {
  CU(h, cudaMemcpyAsync(h->d_scalars_red.p, h->W.scalars, S_TOTAL * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
  if (h->world > 1) {
    int rc = nccl_api().AllReduce(h->W.scalars + 0, h->d_scalars_red.p + 0, 8, kNcclFloat64, kNcclSum, h->comm, h->st);
    if (rc == 0)
      rc = nccl_api().AllReduce(h->W.scalars + 8, h->d_scalars_red.p + 8, 4, kNcclFloat64, kNcclMax, h->comm, h->st);
    if (rc != 0) return h->fail(DBA_ERR_NCCL, "ncclAllReduce: %s", nccl_api().GetErrorString(rc));
  }
  CU(h, cudaMemcpyAsync(h->h_scalars, h->d_scalars_red.p, S_TOTAL * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  CU(h, cudaStreamSynchronize(h->st));
  return DBA_OK;
}

double bytes_per_obs_planes(const dba_handle* h, int planes) { return 16.0 * planes * static_cast<double>(h->n_obs); }

// ---- peer windows (multi-GPU fused PCG tail) ------------------------------------------
// Collective over the ranks of the handle (called from dba_problem_set, which already is): every
// rank allocates its window, the 64-byte CUDA IPC handles travel through one ncclAllGather, every
// rank maps the windows of its peers.  Any failure on any rank (no peer access, IPC disabled in
// the container) leaves ALL ranks on the ncclAllReduce path — agreed through one allreduce.
void close_peer_windows(dba_handle* h) {
  for (int r = 0; r < kMaxPeers; ++r) {
    if (h->win_peer[r]) cudaIpcCloseMemHandle(h->win_peer[r]);
    h->win_peer[r] = nullptr;
  }
  h->p2p_ready = false;
}

int setup_peer_windows(dba_handle* h, size_t nvec) {
  if (h->world == 1) return DBA_OK;
  const char* env = std::getenv("DBA_P2P");
  const bool wanted = h->world <= kMaxPeers && !(env && std::strcmp(env, "0") == 0);
  const size_t slot_len = (std::max<size_t>(nvec, 1) + 31) / 32 * 32;
  const size_t flag_bytes = 256;
  if (wanted && h->p2p_ready && h->win_slot_cap >= slot_len) {  // same decision on every rank (same nvec)
    h->pw.slot_len = static_cast<long long>(slot_len);
    return DBA_OK;
  }
  // tear down the old mapping everywhere before anyone frees its window
  close_peer_windows(h);
  double ok = wanted ? 1.0 : 0.0;
  DevBuf<double> d_ok;
  CU(h, d_ok.alloc(1));
  auto agree = [&]() -> int {  // min over ranks of `ok` (doubles as a barrier)
    const double neg = -ok;
    CU(h, cudaMemcpyAsync(d_ok.p, &neg, sizeof neg, cudaMemcpyHostToDevice, h->st));
    int rc = nccl_api().AllReduce(d_ok.p, d_ok.p, 1, kNcclFloat64, kNcclMax, h->comm, h->st);
    if (rc != 0) return h->fail(DBA_ERR_NCCL, "ncclAllReduce: %s", nccl_api().GetErrorString(rc));
    double out = 0.0;
    CU(h, cudaMemcpyAsync(&out, d_ok.p, sizeof out, cudaMemcpyDeviceToHost, h->st));
    CU(h, cudaStreamSynchronize(h->st));
    ok = -out;
    return DBA_OK;
  };
  int rc = agree();
  if (rc != DBA_OK) return rc;
  if (h->win_local) cudaFree(h->win_local);
  h->win_local = nullptr;
  h->win_bytes = 0;
  h->win_slot_cap = 0;
  if (ok < 0.5) return DBA_OK;

  const size_t data_bytes = 2 * static_cast<size_t>(h->world) * slot_len * sizeof(uint4);
  const size_t bytes = data_bytes + flag_bytes;
  cudaIpcMemHandle_t mine{};
  if (cudaMalloc(&h->win_local, bytes) != cudaSuccess || cudaMemsetAsync(h->win_local, 0, bytes, h->st) != cudaSuccess ||
      cudaIpcGetMemHandle(&mine, h->win_local) != cudaSuccess) {
    cudaGetLastError();
    ok = 0.0;
  }
  CU(h, ensure(h->d_ipc, sizeof(cudaIpcMemHandle_t) * (static_cast<size_t>(h->world) + 1)));
  unsigned char* d_send = h->d_ipc.p + sizeof(cudaIpcMemHandle_t) * h->world;
  CU(h, cudaMemcpyAsync(d_send, &mine, sizeof mine, cudaMemcpyHostToDevice, h->st));
  rc = nccl_api().AllGather(d_send, h->d_ipc.p, sizeof mine, kNcclUint8, h->comm, h->st);
  if (rc != 0) return h->fail(DBA_ERR_NCCL, "ncclAllGather: %s", nccl_api().GetErrorString(rc));
  std::vector<cudaIpcMemHandle_t> all(h->world);
  CU(h, cudaMemcpyAsync(all.data(), h->d_ipc.p, sizeof mine * h->world, cudaMemcpyDeviceToHost, h->st));
  CU(h, cudaStreamSynchronize(h->st));
  if ((rc = agree()) != DBA_OK) return rc;  // did every rank get a window?
  if (ok > 0.5) {
    for (int r = 0; r < h->world && ok > 0.5; ++r) {
      if (r == h->rank) continue;
      if (cudaIpcOpenMemHandle(&h->win_peer[r], all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        h->win_peer[r] = nullptr;
        ok = 0.0;
      }
    }
  }
  if ((rc = agree()) != DBA_OK) return rc;  // did every rank map every peer?
  if (ok < 0.5) {
    close_peer_windows(h);
    if (h->verbose && h->rank == 0) std::fprintf(stderr, "[dba] peer windows unavailable: PCG uses ncclAllReduce\n");
    return DBA_OK;
  }
  h->win_bytes = bytes;
  h->win_slot_cap = slot_len;
  PeerWin& pw = h->pw;
  pw = PeerWin{};
  pw.world = h->world;
  pw.rank = h->rank;
  pw.slot_len = static_cast<long long>(slot_len);
  pw.timeout_ns = 20LL * 1000 * 1000 * 1000;
  if (const char* t = std::getenv("DBA_P2P_TIMEOUT_MS")) pw.timeout_ns = std::max(1LL, std::atoll(t)) * 1000 * 1000;
  for (int r = 0; r < h->world; ++r) {
    char* base = static_cast<char*>(r == h->rank ? h->win_local : h->win_peer[r]);
    pw.ll[r] = reinterpret_cast<uint4*>(base);
  }
  h->p2p_ready = true;
  return DBA_OK;
}

// ---- pieces of dba_problem_set shared by the host build and the device build (ba_build.cu)
struct BuildSizes {
  int64_t nl = 0, ld = 64, n_entries = 0;
  int n_pts = 0, n_tiles = 0, tile_cap = 256, n_chunks = 0, n_partials = 0;
  size_t n_cols = 0;
  int mf_w = 2, two = 0, cb = 0, intr_is_pose = 1;
  bool dense_ok = false;
  int n_dn_batches = 0;
};

// every device buffer of the problem at its size (grow only), solver switches from the environment, peer
// windows, dense-solver work space; nothing here depends on where the index arrays were built
int ensure_work_buffers(dba_handle* h, const BuildSizes& z) {
  const int64_t nl = z.nl, n_entries = z.n_entries;
  const int n_pts = z.n_pts, n_tiles = z.n_tiles, n_chunks = z.n_chunks, n_partials = z.n_partials, two = z.two, cb = z.cb;
  const int n_ext = h->n_ext, n_intr = h->n_intr;
  h->plane_w = 4 + cb + ((two && cb) ? 6 : 0);
  h->j_planes = h->plane_w;
  CU(h, ensure(h->d_obs_xy, nl));
  CU(h, ensure(h->d_obs_ip, nl));
  CU(h, ensure(h->d_obs_ab, nl));
  CU(h, ensure(h->d_obs_lp, nl));
  CU(h, ensure(h->d_tile_obs, n_tiles + 1));
  CU(h, ensure(h->d_tile_pt, n_tiles + 1));
  CU(h, ensure(h->d_pt_first, n_pts + 1));
  CU(h, ensure(h->d_cam_entries, n_entries));
  CU(h, ensure(h->d_cam_chunks, n_chunks));
  CU(h, ensure(h->d_cam_chunk_first, n_ext + 1));
  CU(h, ensure(h->d_tile_meta, n_tiles));
  CU(h, ensure(h->d_part_first_rel, n_entries + n_tiles + 1));
  CU(h, ensure(h->d_items, n_entries));
  CU(h, ensure(h->d_cam_part_first, n_ext + 1));
  CU(h, ensure(h->d_part_dst, n_partials));
  CU(h, ensure(h->d_part_blk, n_partials));
  CU(h, ensure(h->d_obs_lc, nl));
  CU(h, ensure(h->d_mf_cols, std::max<size_t>(z.n_cols * z.mf_w, 1)));
  CU(h, ensure(h->d_items_mf, (two && cb) ? n_entries : 1));
  CU(h, ensure(h->d_part_first, n_entries + n_tiles + 1));
  CU(h, ensure(h->d_mf_rows, static_cast<size_t>(n_ext) * mf_row_len(cb)));
  CU(h, ensure(h->d_mf_T, static_cast<size_t>(n_ext) * (9 + cb)));
  {
    const char* env = std::getenv("DBA_SPMV");
    // default: matrix-free for single-pose problems; composed two-pose rigs (two row fetches, 128
    // registers) measured faster on the plane product (arc1m: 360 vs 304 LM it/s)
    h->mf = two ? 0 : 1;
    if (env && std::strcmp(env, "planes") == 0) h->mf = 0;
    if (env && std::strcmp(env, "mf") == 0) h->mf = 1;
    const char* ef = std::getenv("DBA_PCG_FUSED");
    h->fuse_pcg = !(ef && std::strcmp(ef, "0") == 0);
    const char* et = std::getenv("DBA_MF_TAIL");
    h->mf_tail = !(et && std::strcmp(et, "0") == 0);
    const char* efr = std::getenv("DBA_MF_FRONT");
    h->mf_front = !(efr && std::strcmp(efr, "0") == 0);
    const char* es = std::getenv("DBA_SPECULATE");
    h->speculate = !(es && std::strcmp(es, "0") == 0);
  }
  CU(h, ensure(h->d_partials_q, static_cast<size_t>(n_partials) * std::max(cb, 1)));
  CU(h, ensure(h->d_J, z.ld * h->j_planes));
  CU(h, ensure(h->d_ext_const, n_ext));
  CU(h, ensure(h->d_center, 2 * n_intr));
  CU(h, ensure(h->d_nf, n_intr));
  CU(h, ensure(h->d_nd, n_intr));
  for (int s = 0; s < 3; ++s) {
    CU(h, ensure(h->d_pts[s], 3 * static_cast<size_t>(n_pts)));
    CU(h, ensure(h->d_rot[s], 3 * n_ext));
    CU(h, ensure(h->d_trans[s], 3 * n_ext));
    CU(h, ensure(h->d_focal[s], 2 * n_intr));
    CU(h, ensure(h->d_dist[s], 2 * n_intr));
  }
  for (int s = 0; s < 2; ++s) {
    CU(h, ensure(h->d_pose_rows[s], n_ext));
    CU(h, ensure(h->d_intr_rows[s], n_intr));
  }
  const size_t nvec = static_cast<size_t>(n_ext) * std::max(cb, 1);
  CU(h, ensure(h->d_sp, 3 * static_cast<size_t>(n_pts)));
  CU(h, ensure(h->d_cinv, 6 * static_cast<size_t>(n_pts)));
  CU(h, ensure(h->d_tp, 4 * static_cast<size_t>(n_pts)));
  CU(h, ensure(h->d_dp, 3 * static_cast<size_t>(n_pts)));
  CU(h, ensure(h->d_sc, nvec));
  CU(h, ensure(h->d_cam_acc, nvec * (std::max(cb, 1) + 3)));
  CU(h, ensure(h->d_cam_chunk_acc, static_cast<size_t>(std::max(n_chunks, 1)) * (std::max(cb, 1) * (std::max(cb, 1) + 1) / 2 + 3 * std::max(cb, 1))));
  CU(h, ensure(h->d_minv, nvec * std::max(cb, 1)));
  CU(h, ensure(h->d_dc2, nvec));
  CU(h, ensure(h->d_x, nvec));
  CU(h, ensure(h->d_r, nvec));
  CU(h, ensure(h->d_z, nvec));
  CU(h, ensure(h->d_p, nvec));
  CU(h, ensure(h->d_q, nvec));
  h->q_split = (n_ext >= 296 || !cb) ? 1 : std::min(32, (592 + std::max(n_ext, 1) - 1) / std::max(n_ext, 1));
  if (h->world > 1 && cb) {
    int rc = setup_peer_windows(h, nvec);
    if (rc != DBA_OK) return rc;
  }
  h->dense_ok = z.dense_ok;
  h->Q = DenseWork{};
  if (z.dense_ok) {
    h->Q.n_batches = z.n_dn_batches;
    h->Q.n_pairs = n_ext * (n_ext + 1) / 2;
    CU(h, ensure(h->d_dn_batch, static_cast<size_t>(z.n_dn_batches) + 1));
    CU(h, ensure(h->d_dn_S, nvec * nvec));
    CU(h, ensure(h->d_dn_Spart, static_cast<size_t>(dense_slices(h->Q)) * h->Q.n_pairs * cb * cb));
    h->Q.batch_pt = h->d_dn_batch.p;
    h->Q.S = h->d_dn_S.p;
    h->Q.S_part = h->d_dn_Spart.p;
  }
  CU(h, ensure(h->d_q_split, nvec * static_cast<size_t>(h->q_split)));
  CU(h, ensure(h->d_vec_partials, nvec / 128 + 2 * static_cast<size_t>(n_ext) + 8192));  // k_partials_to_q: one partial per block
  CU(h, ensure(h->d_counters, 4));
  // largest user: apply_step_and_evaluate keeps three ranges side by side (tiles | point update | cost)
  const size_t n_part = static_cast<size_t>((nl + 255) / 256) + 3 * static_cast<size_t>(n_tiles) + 3 +
                        2 * static_cast<size_t>((3 * static_cast<int64_t>(n_pts) + 255) / 256) + 256;
  CU(h, ensure(h->d_partA, n_part));
  CU(h, ensure(h->d_partB, 3 * static_cast<size_t>((std::max(n_ext, n_intr) + 5) / 6) + 64));  // k_camera_finalize: 7 or 10 blocks per CTA
  CU(h, ensure(h->d_scalars, S_TOTAL));
  CU(h, ensure(h->d_scalars_red, S_TOTAL));
  CU(h, ensure(h->d_pcg_scal, 8));
  CU(h, ensure(h->d_pcg_state, 4));
  CU(h, cudaMemsetAsync(h->d_counters.p, 0, 4 * sizeof(unsigned int), h->st));
  CU(h, cudaMemsetAsync(h->d_scalars.p, 0, S_TOTAL * sizeof(double), h->st));
  CU(h, cudaMemsetAsync(h->d_x.p, 0, std::max<size_t>(nvec, 1) * sizeof(double), h->st));
  CU(h, cudaMemsetAsync(h->d_pcg_state.p, 0, 4 * sizeof(int), h->st));
  if (h->world > 1) CU(h, ensure(h->d_full_pts, 3 * static_cast<size_t>(h->n_pts_global)));  // dba_params_get
  return DBA_OK;
}

// camera-side parameters and constant tables of the caller (slot 2 = pristine copy for dba_params_reset);
// also what dba_problem_update needs to keep
int upload_params(dba_handle* h, const dba_problem* p) {
  const int n_ext = h->n_ext, n_intr = h->n_intr;
  auto up = [&](void* dst, const void* src, size_t bytes) -> cudaError_t {
    if (bytes == 0) return cudaSuccess;
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->st);
  };
  std::vector<uint8_t>& ext_const = h->keep.ext_const;
  ext_const.assign(std::max(n_ext, 1), 0);
  h->any_const = false;
  if (p->ext_const)
    for (int i = 0; i < n_ext; ++i) {
      ext_const[i] = p->ext_const[i] ? 1 : 0;
      h->any_const |= ext_const[i] != 0;
    }
  CU(h, up(h->d_ext_const.p, ext_const.data(), n_ext));
  CU(h, up(h->d_center.p, p->intr_center, 2 * sizeof(double) * n_intr));
  CU(h, up(h->d_nf.p, p->intr_nf, sizeof(int) * n_intr));
  CU(h, up(h->d_nd.p, p->intr_nd, sizeof(int) * n_intr));
  CU(h, up(h->d_rot[2].p, p->ext_rot, 3 * sizeof(double) * n_ext));
  CU(h, up(h->d_trans[2].p, p->ext_trans, 3 * sizeof(double) * n_ext));
  CU(h, up(h->d_focal[2].p, p->intr_focal, 2 * sizeof(double) * n_intr));
  CU(h, up(h->d_dist[2].p, p->intr_dist, 2 * sizeof(double) * n_intr));
  h->keep.center.assign(p->intr_center, p->intr_center + 2 * static_cast<size_t>(n_intr));
  h->keep.nf.assign(p->intr_nf, p->intr_nf + n_intr);
  h->keep.nd.assign(p->intr_nd, p->intr_nd + n_intr);
  h->keep.free_intrinsics = p->free_intrinsics;
  return DBA_OK;
}

// DeviceProblem / ParamSet / WorkArrays point at the buffers
int bind_problem(dba_handle* h, const BuildSizes& z) {
  h->n_cam_entries = z.n_entries;
  DeviceProblem& D = h->D;
  D.n_obs = z.nl;
  D.ld = z.ld;
  D.n_pts = z.n_pts;
  D.n_ext = h->n_ext;
  D.n_intr = h->n_intr;
  D.n_tiles = z.n_tiles;
  D.tile = z.tile_cap;
  D.cb = z.cb;
  D.two = z.two;
  D.n_blocks = h->n_ext;
  D.obs_xy = h->d_obs_xy.p;
  D.obs_ip = h->d_obs_ip.p;
  D.tile_obs = h->d_tile_obs.p;
  D.tile_pt = h->d_tile_pt.p;
  D.pt_first = h->d_pt_first.p;
  D.cam_entries = h->d_cam_entries.p;
  D.cam_chunks = h->d_cam_chunks.p;
  D.cam_chunk_first = h->d_cam_chunk_first.p;
  D.n_chunks = z.n_chunks;
  D.J = h->d_J.p;
  D.tile_meta = h->d_tile_meta.p;
  D.obs_ab = h->d_obs_ab.p;
  D.obs_lp = h->d_obs_lp.p;
  D.part_first_rel = h->d_part_first_rel.p;
  D.items = h->d_items.p;
  D.cam_part_first = h->d_cam_part_first.p;
  D.n_partials = z.n_partials;
  D.mf_cols = h->d_mf_cols.p;
  D.items_mf = h->d_items_mf.p;
  D.part_dst = h->d_part_dst.p;
  D.part_blk = h->d_part_blk.p;
  D.obs_lc = h->d_obs_lc.p;
  D.intr_is_pose = z.intr_is_pose;
  D.part_first = h->d_part_first.p;
  for (int s = 0; s < 2; ++s) {
    ParamSet& P = h->P[s];
    P.pts = h->d_pts[s].p;
    P.ext_rot = h->d_rot[s].p;
    P.ext_trans = h->d_trans[s].p;
    P.focal = h->d_focal[s].p;
    P.dist = h->d_dist[s].p;
    P.center = h->d_center.p;
    P.nf = h->d_nf.p;
    P.nd = h->d_nd.p;
    P.pose_rows = h->d_pose_rows[s].p;
    P.intr_rows = h->d_intr_rows[s].p;
  }
  WorkArrays& W = h->W;
  W.sp = h->d_sp.p;
  W.sc = h->d_sc.p;
  W.cinv = h->d_cinv.p;
  W.tp = h->d_tp.p;
  W.dp = h->d_dp.p;
  W.cam_acc = h->d_cam_acc.p;
  W.cam_chunk_acc = h->d_cam_chunk_acc.p;
  W.minv = h->d_minv.p;
  W.dc2 = h->d_dc2.p;
  W.x = h->d_x.p;
  W.r = h->d_r.p;
  W.z = h->d_z.p;
  W.p = h->d_p.p;
  W.q = h->d_q.p;
  W.scalars = h->d_scalars.p;
  W.pcg_state = h->d_pcg_state.p;
  W.pcg_scal = h->d_pcg_scal.p;
  W.partials_q = h->d_partials_q.p;
  W.q_split = h->d_q_split.p;
  W.mf_rows = h->d_mf_rows.p;
  W.mf_T = h->d_mf_T.p;
  W.vec_partials = h->d_vec_partials.p;
  W.counters = h->d_counters.p;
  W.trace = nullptr;
  if (std::getenv("DBA_TAIL_TRACE")) {
    CU(h, ensure(h->d_trace, 16));
    unsigned long long init[16] = {0};
    init[10] = ~0ull;
    CU(h, cudaMemcpy(h->d_trace.p, init, sizeof init, cudaMemcpyHostToDevice));
    W.trace = h->d_trace.p;
  }
  h->Q.fail_flag = h->d_pcg_state.p + 3;
  return DBA_OK;
}

// ---- LM building blocks ---------------------------------------------------------------
void fill_ones(dba_handle* h, double* p, size_t n) {
  std::vector<double> ones(n, 1.0);
  cudaMemcpyAsync(p, ones.data(), n * sizeof(double), cudaMemcpyHostToDevice, h->st);
  cudaStreamSynchronize(h->st);
}

// Jacobian (+cost) at the current parameters.  first == true also establishes the Jacobi scales.
int evaluate_jacobian(dba_handle* h, bool first, bool jacobi_scaling) {
  const DeviceProblem& D = h->D;
  const ParamSet& P = h->P[h->cur];
  const int nplanes = h->planeless ? 4 : 4 + h->cb + (h->two && h->cb ? 6 : 0);
  // read xy(16)+idx(8), write r + Jp + Jc planes (no Jc planes when every consumer recomputes F)
  const double k1_bytes = (24.0 + 16.0 * nplanes) * static_cast<double>(h->n_obs);
  {
    Scope s(h, "pose_rows");
    launch_pose_rows(P, h->d_ext_const.p, h->freeze, h->n_ext, h->n_intr, h->st);
  }
  if (first) {
    if (jacobi_scaling) {
      {
        Scope s(h, "jacobian", k1_bytes);
        launch_jacobian(D, P, h->W, h->planeless ? 0 : h->cb, h->planeless ? 0 : h->two, /*unit_scale=*/1, nullptr, h->st);
      }
      {
        Scope s(h, "point_prepare", bytes_per_obs_planes(h, 4));
        launch_point_prepare(D, h->W, 1.0, 0.0, 0.0, /*mode=*/0, h->d_partA.p, h->st);
      }
      if (h->cb) {
        {
          Scope s(h, "camera_gather", 0.0, 2);
          launch_camera_gather(D, h->W, 0, h->st, h->mf_front ? &P : nullptr);
        }
        int rc = allreduce(h, h->W.cam_acc, static_cast<size_t>(D.n_blocks) * h->cb * (h->cb + 3), kNcclSum);
        if (rc != DBA_OK) return rc;
        Scope s(h, "camera_scales");
        launch_camera_scales(D, h->W, h->st);
      }
    } else {
      fill_ones(h, h->W.sp, 3 * static_cast<size_t>(h->n_pts));
      if (h->cb) fill_ones(h, h->W.sc, static_cast<size_t>(D.n_blocks) * h->cb);
    }
  }
  {
    Scope s(h, "jacobian", k1_bytes);
    launch_jacobian(D, P, h->W, h->planeless ? 0 : h->cb, h->planeless ? 0 : h->two, /*unit_scale=*/0, h->d_partA.p, h->st);
  }
  {
    Scope s(h, "reduce");
    launch_reduce_sum(h->d_partA.p, jacobian_partials(D), 1, 0, h->W.scalars + S_COST, h->st);
  }
  if (h->mf && h->cb) {
    Scope s(h, "mf_rows");
    launch_mf_rows(D, P, h->W, h->st);
  }
  CU(h, cudaGetLastError());
  return DBA_OK;
}

// Schur front half for the given radius; leaves gradient norms / failure flags in the scalars.
int prepare_step(dba_handle* h, double radius, const dba_solve_options& o) {
  const DeviceProblem& D = h->D;
  {
    Scope s(h, "point_prepare", bytes_per_obs_planes(h, 4));
    launch_point_prepare(D, h->W, radius, o.min_lm_diagonal, o.max_lm_diagonal, 1, h->d_partA.p, h->st);
  }
  {
    Scope s(h, "reduce");
    ReduceJobs j;
    j.sum(h->d_partA.p, D.n_tiles, 3, 0, h->W.scalars + S_GSQ_PT);
    j.max(h->d_partA.p, D.n_tiles, 3, 1, h->W.scalars + S_GMAX_PT);
    j.sum(h->d_partA.p, D.n_tiles, 3, 2, h->W.scalars + S_BAD_PT);
    launch_reduce_multi(j, h->st);
  }
  if (h->cb) {
    {
      // gather: Jc + Jp + r planes at sector granularity, plus C^-1 and t per observation
      // (the dense reduced system does not need the block-Jacobi blocks: mode 2)
      Scope s(h, "camera_gather", 0.0, 2);
      launch_camera_gather(D, h->W, h->use_dense ? 2 : 1, h->st, h->mf_front ? &h->P[h->cur] : nullptr);
    }
    int rc = allreduce(h, h->W.cam_acc, static_cast<size_t>(D.n_blocks) * h->cb * (h->cb + 3), kNcclSum);
    if (rc != DBA_OK) return rc;
    {
      Scope s(h, "camera_finalize");
      launch_camera_finalize(D, h->W, radius, o.min_lm_diagonal, o.max_lm_diagonal, h->d_partB.p, h->use_dense ? 0 : 1, h->st);
    }
    const int g = camera_finalize_grid(D);
    Scope s(h, "reduce");
    ReduceJobs j;
    j.sum(h->d_partB.p, g, 3, 0, h->W.scalars + S_GSQ_CAM);
    j.max(h->d_partB.p, g, 3, 1, h->W.scalars + S_GMAX_CAM);
    j.sum(h->d_partB.p, g, 3, 2, h->W.scalars + S_BAD_CAM);
    launch_reduce_multi(j, h->st);
  } else {
    CU(h, cudaMemsetAsync(h->W.scalars + S_GSQ_CAM, 0, 3 * sizeof(double), h->st));
  }
  CU(h, cudaGetLastError());
  return DBA_OK;
}

// Block-Jacobi PCG on the implicit Schur complement; returns iterations run.
int pcg_solve(dba_handle* h, const dba_solve_options& o, int* iters_out) {
  const DeviceProblem& D = h->D;
  const int nplanes = 3 + h->cb + (h->two ? 6 : 0);
  // one pass over the planes (indices 8 B + Jp, Jc planes) + the tile partials written once
  // and read once by the per-camera sum
  const double tile_bytes = (8.0 + 16.0 * nplanes) * static_cast<double>(h->n_obs) +
                            8.0 * h->cb * static_cast<double>(D.n_partials);
  const double part_bytes = 8.0 * h->cb * static_cast<double>(D.n_partials);
  // matrix-free product: column indices + per-point X, sp, C^-1, segment bounds + partials written once
  const double mf_bytes = (h->cb == 9 ? 8.0 : 16.0) * static_cast<double>(D.n_tiles) * D.tile + 100.0 * static_cast<double>(h->n_pts) +
                          (8.0 * h->cb + 4.0) * static_cast<double>(D.n_partials);
  {
    Scope s(h, "pcg_init");
    launch_pcg_init(D, h->W, h->st);
  }
  if (h->mf) {
    Scope s(h, "pcg_vector");
    launch_mf_direction(D, h->W, /*init=*/1, h->st);
  }
  const double tol2 = o.pcg_rel_tolerance * o.pcg_rel_tolerance;
  const int max_it = std::max(0, o.pcg_max_iterations);
  const int nvec = D.n_blocks * h->cb;
  int issued = 0;
  const int check_every = (o.pcg_rel_tolerance > 0.0) ? 8 : max_it;
  // q sum (+ the cross-rank exchange through the peer windows) + D^2 p + p.q + the vector phases in one kernel
  const int fused = (h->world == 1 || h->p2p_ready) ? 1 : 0;
  const PeerWin no_peers{};
  *iters_out = 0;
  while (issued < max_it) {
    const int batch = std::min(check_every > 0 ? check_every : max_it, max_it - issued);
    for (int i = 0; i < batch; ++i) {
      if (h->mf) {
        // default: the rest of the PCG iteration runs as the epilogue of the product kernel
        MfTail tail;
        if (fused && h->fuse_pcg && h->mf_tail) {
          tail.fuse = std::max(h->q_split, 1);
          tail.tol2 = tol2;
          tail.min_iter = o.pcg_min_iterations;
          if (h->world > 1) {
            h->pw.seq = ++h->p2p_seq;
            tail.pw = h->pw;
          }
        }
        Scope s(h, tail.fuse ? "spmv_mf_pcg" : "spmv_mf", mf_bytes + (tail.fuse ? part_bytes : 0.0));
        if (launch_spmv_mf(D, h->P[h->cur], h->W, tail, h->st) != 0)
          return h->fail(DBA_ERR_CUDA, "cooperative launch of k_spmv_mf failed: %s", cudaGetErrorString(cudaGetLastError()));
        if (tail.fuse) continue;
      } else {
        Scope s(h, "spmv_tile", tile_bytes);
        launch_spmv_tile(D, h->W, h->st);
      }
      if (fused && h->fuse_pcg) {
        Scope s(h, "pcg_fused", part_bytes);
        if (h->world > 1) h->pw.seq = ++h->p2p_seq;
        if (launch_pcg_fused(D, h->W, h->mf, tol2, o.pcg_min_iterations, h->q_split, h->world > 1 ? h->pw : no_peers, h->st) != 0)
          return h->fail(DBA_ERR_CUDA, "cooperative launch of k_pcg_fused failed: %s", cudaGetErrorString(cudaGetLastError()));
        continue;
      }
      const int fuse_dot = (h->world == 1 && h->q_split == 1) ? 1 : 0;
      {
        Scope s(h, "partials_to_q", part_bytes);
        launch_partials_to_q(D, h->W, fuse_dot, h->q_split, h->mf, h->st);
      }
      if (h->world > 1) {
        if (h->q_split > 1) {  // fold the slices before the allreduce
          Scope s(h, "pcg_vector");
          launch_fold_q(D, h->W, h->q_split, h->st);
        }
        int rc = allreduce(h, h->W.q, nvec, kNcclSum);
        if (rc != DBA_OK) return rc;
      }
      Scope s(h, "pcg_vector", 0.0, fuse_dot ? 2 : 3);
      if (!fuse_dot) launch_pcg_dot(D, h->W, h->world > 1 ? 1 : h->q_split, h->st);
      launch_pcg_step(D, h->W, tol2, o.pcg_min_iterations, h->st);
      if (h->mf)
        launch_mf_direction(D, h->W, /*init=*/0, h->st);
      else
        launch_pcg_direction(D, h->W, h->st);
    }
    issued += batch;
    CU(h, cudaMemcpyAsync(h->h_pcg_state, h->W.pcg_state, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->st));
    if (o.pcg_rel_tolerance <= 0.0) {
      // fixed iteration count: nothing to decide here, the state is read together with the LM
      // scalars at the next synchronisation (pcg_collect) — one host round trip less per LM iteration
      h->pcg_pending = true;
      *iters_out = issued;
      break;
    }
    CU(h, cudaStreamSynchronize(h->st));
    *iters_out = h->h_pcg_state[0];
    if (h->h_pcg_state[2]) return h->fail(DBA_ERR_NCCL, "peer exchange timed out: a rank of this handle stopped responding");
    if (h->h_pcg_state[1]) break;
    if (issued >= max_it) h->pcg_unconverged++;  // the tolerance was asked for and not reached
  }
  if (max_it == 0) {
    CU(h, cudaMemcpyAsync(h->h_pcg_state, h->W.pcg_state, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CU(h, cudaStreamSynchronize(h->st));
  }
  CU(h, cudaGetLastError());
  return DBA_OK;
}

// DENSE_SCHUR (reference sfm.cc:67): explicit reduced system S from the planes, allreduce over the
// ranks, Cholesky + substitutions in one CTA -> W.x.  A failed factorisation leaves NaN in W.x (the
// step is then invalid, as with the CPU oracle) and raises pcg_state[3], read at the next
// synchronisation by pcg_collect.
int dense_solve(dba_handle* h, int* iters_out) {
  const DeviceProblem& D = h->D;
  const int n = D.n_blocks * h->cb;
  const int nplanes = 3 + h->cb + (h->two ? 6 : 0);
  CU(h, cudaMemsetAsync(h->W.pcg_state, 0, 4 * sizeof(int), h->st));
  if (!h->Q.Z && h->Q.n_batches > 0) {  // the Z image exists only for handles that solve densely
    CU(h, ensure(h->d_dn_Z, static_cast<size_t>(h->Q.n_batches) * kDnEntCap * 3 * h->cb));
    CU(h, ensure(h->d_dn_Zent, static_cast<size_t>(h->Q.n_batches) * kDnEntCap));
    CU(h, ensure(h->d_dn_Zcount, static_cast<size_t>(h->Q.n_batches)));
    h->Q.Z = h->d_dn_Z.p;
    h->Q.Zent = h->d_dn_Zent.p;
    h->Q.Zcount = h->d_dn_Zcount.p;
  }
  {
    Scope s(h, "schur_dense", 16.0 * nplanes * static_cast<double>(h->n_obs), h->Q.n_pair_chunks > 0 ? 4 : 3);
    if (launch_schur_dense(D, h->W, h->Q, /*add_diag=*/h->rank == 0 ? 1 : 0, h->st) != 0) return h->fail(DBA_ERR_UNSUPPORTED, "dense reduced system: unsupported camera block");
  }
  int rc = allreduce(h, h->Q.S, static_cast<size_t>(n) * n, kNcclSum);
  if (rc != DBA_OK) return rc;
  {
    Scope s(h, "dense_cholesky", 8.0 * n * n);
    launch_dense_cholesky(D, h->W, h->Q, h->st);
  }
  CU(h, cudaMemcpyAsync(h->h_pcg_state, h->W.pcg_state, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->st));
  h->pcg_pending = true;
  *iters_out = 1;
  CU(h, cudaGetLastError());
  return DBA_OK;
}

// after a stream synchronisation: the linear-solver state copied by a pcg_solve / dense_solve that
// did not wait for it
int pcg_collect(dba_handle* h, int* iters) {
  if (!h->pcg_pending) return DBA_OK;
  h->pcg_pending = false;
  if (h->use_dense) {
    *iters = 1;
    if (h->h_pcg_state[3]) h->dense_failures++;
    return DBA_OK;
  }
  *iters = h->h_pcg_state[0];
  if (h->h_pcg_state[2]) return h->fail(DBA_ERR_NCCL, "peer exchange timed out: a rank of this handle stopped responding");
  return DBA_OK;
}

// step -> candidate parameters, model cost change, norms, candidate cost.
// speculate: evaluate the candidate with the Jacobian kernel instead of the residual-only kernel
// (same residual code, same partial-sum order), so that an accepted step needs no second pass over
// the observations: the planes, the cost and the matrix-free camera rows are already those of the
// new point.  A step that is not accepted restores them (restore_current_jacobian).
int apply_step_and_evaluate(dba_handle* h, bool speculate) {
  const DeviceProblem& D = h->D;
  const ParamSet& cur = h->P[h->cur];
  const ParamSet& cand = h->P[1 - h->cur];
  const int nplanes = h->planeless ? 4 : 4 + h->cb + (h->two && h->cb ? 6 : 0);
  const int gp = update_points_grid(D), gc = update_cameras_grid(D);
  // three disjoint partial ranges of d_partA so that ONE launch reduces everything at the end
  double* pa_model = h->d_partA.p;
  double* pa_upd = pa_model + ((D.n_tiles + 31) / 32) * 32;
  double* pa_cost = pa_upd + ((2 * gp + 31) / 32) * 32;
  {
    Scope s(h, "back_substitute", (8.0 + 16.0 * nplanes) * static_cast<double>(h->n_obs));
    launch_back_substitute(D, h->W, pa_model, h->st, h->planeless ? &cur : nullptr);
  }
  {
    Scope s(h, "param_update", 0.0, 2);
    launch_update_points(D, cur, cand, h->W, pa_upd, h->st);
    launch_update_cameras(D, cur, cand, h->W, h->d_partB.p, h->st);
  }
  {
    Scope s(h, "pose_rows");
    launch_pose_rows(cand, h->d_ext_const.p, h->freeze, h->n_ext, h->n_intr, h->st);
  }
  if (speculate) {
    {
      Scope s(h, "jacobian", (24.0 + 16.0 * nplanes) * static_cast<double>(h->n_obs));
      launch_jacobian(D, cand, h->W, h->planeless ? 0 : h->cb, h->planeless ? 0 : h->two, /*unit_scale=*/0, pa_cost, h->st);
    }
    if (h->mf && h->cb) {
      Scope s(h, "mf_rows");
      launch_mf_rows(D, cand, h->W, h->st);
    }
    h->jacobian_at_candidate = true;
  } else {
    Scope s(h, "cost", 24.0 * static_cast<double>(h->n_obs));
    launch_cost(D, cand, pa_cost, nullptr, h->st);
  }
  {
    Scope s(h, "reduce");
    ReduceJobs j;
    j.sum(pa_model, D.n_tiles, 1, 0, h->W.scalars + S_MODEL);
    j.sum(pa_upd, gp, 2, 0, h->W.scalars + S_STEP_PT);
    j.sum(pa_upd, gp, 2, 1, h->W.scalars + S_X_PT);
    j.sum(h->d_partB.p, gc, 2, 0, h->W.scalars + S_STEP_CAM);
    j.sum(h->d_partB.p, gc, 2, 1, h->W.scalars + S_X_CAM);
    j.sum(pa_cost, speculate ? jacobian_partials(D) : cost_grid(D), 1, 0, h->W.scalars + S_CAND);
    launch_reduce_multi(j, h->st);
  }
  CU(h, cudaGetLastError());
  return DBA_OK;
}

// The speculative evaluation left the Jacobian of a candidate that was not accepted: back to x.
int restore_current_jacobian(dba_handle* h, bool jacobi_scaling) {
  if (!h->jacobian_at_candidate) return DBA_OK;
  h->jacobian_at_candidate = false;
  return evaluate_jacobian(h, /*first=*/false, jacobi_scaling);
}

// The candidate of a speculative evaluation was accepted (h->cur already flipped): its cost is the
// cost at x now; planes and camera rows are in place.
int adopt_candidate_jacobian(dba_handle* h) {
  h->jacobian_at_candidate = false;
  CU(h, cudaMemcpyAsync(h->W.scalars + S_COST, h->W.scalars + S_CAND, sizeof(double), cudaMemcpyDeviceToDevice, h->st));
  return DBA_OK;
}

void print_progress_header() {
  std::printf(
      "iter      cost      cost_change  |gradient|   |step|    tr_ratio  tr_radius  ls_iter  iter_time  total_time\n");
}

}  // namespace

// ======================================================================== C ABI
extern "C" {

int dba_abi_version(void) { return DBA_ABI_VERSION; }

int dba_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return DBA_ERR_NO_DEVICE;
  }
  return n;
}

int dba_nccl_unique_id(void* out128) {
  if (!out128) return DBA_ERR_INVALID_ARGUMENT;
  if (!nccl_api().load()) return DBA_ERR_NCCL;
  NcclUniqueId id;
  if (nccl_api().GetUniqueId(&id) != 0) return DBA_ERR_NCCL;
  std::memcpy(out128, &id, sizeof id);
  return DBA_OK;
}

const char* dba_last_error(const dba_handle* h) { return h ? h->error.c_str() : g_create_error.c_str(); }

int dba_create(dba_handle** out, const dba_config* cfg) {
  if (!out || !cfg) {
    g_create_error = "dba_create: null argument";
    return DBA_ERR_INVALID_ARGUMENT;
  }
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    g_create_error = "no CUDA device available: this engine has no CPU path";
    return DBA_ERR_NO_DEVICE;
  }
  if (cfg->device < 0 || cfg->device >= ndev || cfg->world_size < 1 || cfg->rank < 0 || cfg->rank >= cfg->world_size) {
    g_create_error = "dba_create: bad device / rank / world_size";
    return DBA_ERR_INVALID_ARGUMENT;
  }
  dba_handle* h = new dba_handle;
  h->device = cfg->device;
  h->rank = cfg->rank;
  h->world = cfg->world_size;
  h->verbose = cfg->verbose;
  cudaError_t e = cudaSetDevice(h->device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&h->h_scalars), S_TOTAL * sizeof(double));
  if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&h->h_pcg_state), 4 * sizeof(int));
  auto release = [&]() {  // error exits of dba_create: nothing allocated so far may leak
    if (h->h_scalars) cudaFreeHost(h->h_scalars);
    if (h->h_pcg_state) cudaFreeHost(h->h_pcg_state);
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
  };
  if (e != cudaSuccess) {
    g_create_error = std::string("CUDA initialisation failed: ") + cudaGetErrorString(e);
    release();
    return DBA_ERR_CUDA;
  }
  if (h->world > 1) {
    if (!cfg->nccl_unique_id || !nccl_api().load()) {
      g_create_error = "world_size > 1 needs NCCL (libnccl.so.2) and a unique id";
      release();
      return DBA_ERR_NCCL;
    }
    NcclUniqueId id;
    std::memcpy(&id, cfg->nccl_unique_id, sizeof id);
    int rc = nccl_api().CommInitRank(&h->comm, h->world, id, h->rank);
    if (rc != 0) {
      g_create_error = std::string("ncclCommInitRank: ") + nccl_api().GetErrorString(rc);
      release();
      return DBA_ERR_NCCL;
    }
  }
  *out = h;
  return DBA_OK;
}

void dba_destroy(dba_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->st) cudaStreamSynchronize(h->st);
  drain_events(h);
  for (cudaEvent_t e : h->event_pool) cudaEventDestroy(e);
  for (cudaEvent_t e : h->ev_solve)
    if (e) cudaEventDestroy(e);
  close_peer_windows(h);
  if (h->comm && h->win_local && h->d_scalars_red.p) {
    // explicit barrier (dba_destroy is collective): every peer has closed its mapping of this rank's
    // window before the exported memory is freed
    if (nccl_api().AllReduce(h->d_scalars_red.p, h->d_scalars_red.p, 1, kNcclFloat64, kNcclSum, h->comm, h->st) == 0)
      cudaStreamSynchronize(h->st);
  }
  if (h->comm) nccl_api().CommDestroy(h->comm);
  if (h->win_local) cudaFree(h->win_local);
  delete h->upload;
  if (h->h_scalars) cudaFreeHost(h->h_scalars);
  if (h->h_pcg_state) cudaFreeHost(h->h_pcg_state);
  if (h->st) cudaStreamDestroy(h->st);
  delete h;
}

int dba_kernel_stats_enable(dba_handle* h, int32_t enable) {
  if (!h) return DBA_ERR_INVALID_ARGUMENT;
  h->stats_enabled = enable != 0;
  return DBA_OK;
}
int dba_kernel_stats_reset(dba_handle* h) {
  if (!h) return DBA_ERR_INVALID_ARGUMENT;
  drain_events(h);
  for (auto& r : h->records) {
    r.launches = 0;
    r.total_ms = 0.0;
  }
  return DBA_OK;
}
int dba_kernel_stats(dba_handle* h, dba_kernel_stat* out, int32_t capacity) {
  if (!h || (!out && capacity > 0)) return DBA_ERR_INVALID_ARGUMENT;
  cudaSetDevice(h->device);
  drain_events(h);
  int n = 0;
  for (const auto& r : h->records) {
    if (n >= capacity) break;
    std::snprintf(out[n].name, sizeof out[n].name, "%s", r.name.c_str());
    out[n].launches = r.launches;
    out[n].total_ms = r.total_ms;
    out[n].algorithmic_bytes = r.algorithmic_bytes;
    ++n;
  }
  return n;
}

// ------------------------------------------------------------------- sharding plan
int dba_shard_plan(const dba_problem* p, int32_t world_size, int32_t* pt_begin, int64_t* obs_count) {
  if (!p || world_size < 1 || !pt_begin || p->n_obs < 0 || p->n_pts < 0 || (p->n_obs > 0 && !p->obs_pt))
    return DBA_ERR_INVALID_ARGUMENT;
  std::vector<int64_t> first(static_cast<size_t>(p->n_pts) + 1, 0);
  for (int64_t i = 0; i < p->n_obs; ++i) {
    if (p->obs_pt[i] < 0 || p->obs_pt[i] >= p->n_pts) return DBA_ERR_INVALID_ARGUMENT;
    first[p->obs_pt[i] + 1]++;
  }
  for (int i = 0; i < p->n_pts; ++i) first[i + 1] += first[i];
  pt_begin[0] = 0;
  for (int r = 1; r < world_size; ++r) {
    const int64_t target = p->n_obs * r / world_size;
    int cut = static_cast<int>(std::lower_bound(first.begin(), first.end(), target) - first.begin());
    cut = std::min(cut, p->n_pts);
    pt_begin[r] = std::max(cut, pt_begin[r - 1]);
  }
  pt_begin[world_size] = p->n_pts;
  if (obs_count)
    for (int r = 0; r < world_size; ++r) obs_count[r] = first[pt_begin[r + 1]] - first[pt_begin[r]];
  return DBA_OK;
}

// Host threads of the upload: with one rank per GPU on one host, every rank builds its shard at the
// same time, so each takes its share of the cores instead of oversubscribing them world-fold
// (DBA_HOST_THREADS overrides); the caller's OpenMP setting is restored on the way out.
struct OmpThreadScope {
  int saved;
  explicit OmpThreadScope(int world) : saved(omp_get_max_threads()) {
    int n = world > 1 ? std::max(1, omp_get_num_procs() / world) : 0;
    if (const char* e = std::getenv("DBA_HOST_THREADS")) n = std::atoi(e);
    if (n > 0) omp_set_num_threads(n);
  }
  ~OmpThreadScope() { omp_set_num_threads(saved); }
};

// ------------------------------------------------------------------- device-side problem construction
// SURVEY 8 f-3.  Point-sorted, single-pose input only; *handled = 0 leaves everything to the host build.
// With several ranks every rank stages and builds ONLY its shard (the host build hands every rank the whole
// problem and each builds its shard with cores / world threads).
static int problem_set_device(dba_handle* h, const dba_problem* p, int* handled) {
  *handled = 0;
  const int64_t n = p->n_obs;
  const int n_ext = p->n_ext, n_intr = p->n_intr;
  if (n <= 0 || p->n_pts <= 0 || n_ext <= 0) return DBA_OK;
  const bool timing = std::getenv("DBA_TIMING") != nullptr;
  double t_mark = now_s();
  auto mark = [&](const char* what) {
    if (!timing) return;
    const double t = now_s();
    std::fprintf(stderr, "[dba_problem_set/device] %-26s %8.2f ms\n", what, 1e3 * (t - t_mark));
    t_mark = t;
  };
  for (int i = 0; i < n_intr; ++i)
    if (p->intr_nf[i] < 1 || p->intr_nf[i] > 2 || p->intr_nd[i] < 0 || p->intr_nd[i] > 2)
      return h->fail(DBA_ERR_INVALID_ARGUMENT, "intrinsic %d: nf must be 1|2 and nd 0|1|2", i);
  const int freeze = p->freeze_camera != 0;
  const int free_intr = (!freeze && p->free_intrinsics) ? 1 : 0;
  if (free_intr) {
    if (n_ext != n_intr) return h->fail(DBA_ERR_UNSUPPORTED, "free_intrinsics needs one intrinsic per extrinsic");
    for (int i = 0; i < n_intr; ++i)
      if (p->intr_nf[i] != 1 || p->intr_nd[i] != 2) return h->fail(DBA_ERR_UNSUPPORTED, "free_intrinsics needs nf=1, nd=2");
  }
  const int cb = freeze ? 0 : (free_intr ? 9 : 6);
  // shard of this rank, assuming sorted input (verified on the device below): same cuts as dba_shard_plan
  auto cut_pt = [&](int r) -> int {
    if (r <= 0) return 0;
    if (r >= h->world) return p->n_pts;
    const int64_t target = n * r / h->world;
    return target <= 0 ? 0 : std::min(p->obs_pt[target - 1] + 1, p->n_pts);
  };
  auto first_obs_of = [&](int pt) -> int64_t {  // lower bound in the (sorted) point column
    return std::lower_bound(p->obs_pt, p->obs_pt + n, pt) - p->obs_pt;
  };
  int pt_lo = cut_pt(h->rank);
  for (int r = 1; r < h->rank; ++r) pt_lo = std::max(pt_lo, cut_pt(r));
  int pt_hi = std::max(cut_pt(h->rank + 1), pt_lo);
  if (h->rank + 1 == h->world) pt_hi = p->n_pts;
  if (pt_lo < 0 || pt_hi > p->n_pts) return DBA_OK;  // garbage point ids: let the host path report them
  const int64_t obs_lo = h->world > 1 ? first_obs_of(pt_lo) : 0, obs_hi = h->world > 1 ? first_obs_of(pt_hi) : n;
  const int64_t nl = obs_hi - obs_lo;
  const int n_pts = pt_hi - pt_lo;
  if (nl >= (int64_t{1} << 30)) return DBA_OK;
  // ---- stage the raw shard through the pinned arena (all host cores of this rank), one DMA per array
  OmpThreadScope omp_scope(h->world);
  PinnedArena& A = arena_of(h);
  CU(h, cudaStreamSynchronize(h->st));
  const bool have_b = p->obs_pose_b != nullptr;
  CU(h, A.reserve(pad256(nl * 16) + (have_b ? 4 : 3) * pad256(nl * 4) + pad256(static_cast<size_t>(n_pts) * 24) +
                  pad256((static_cast<size_t>(n_pts) + 2) * 4) + 8192));
  double* s_xy = A.take<double>(static_cast<size_t>(std::max<int64_t>(2 * nl, 1)));
  int* s_pt = A.take<int>(static_cast<size_t>(std::max<int64_t>(nl, 1)));
  int* s_a = A.take<int>(static_cast<size_t>(std::max<int64_t>(nl, 1)));
  int* s_b = have_b ? A.take<int>(static_cast<size_t>(std::max<int64_t>(nl, 1))) : nullptr;
  int* s_in = A.take<int>(static_cast<size_t>(std::max<int64_t>(nl, 1)));
  double* s_pts = A.take<double>(static_cast<size_t>(std::max(3 * n_pts, 1)));
  int* s_first = A.take<int>(static_cast<size_t>(n_pts) + 2);
  int* s_flags = A.take<int>(8);
  auto stage = [&](void* dst, const void* src, size_t bytes) {
    const int64_t chunks = static_cast<int64_t>((bytes + (1 << 20) - 1) >> 20);
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < chunks; ++c) {
      const size_t a = static_cast<size_t>(c) << 20, b = std::min(bytes, a + (size_t{1} << 20));
      std::memcpy(static_cast<char*>(dst) + a, static_cast<const char*>(src) + a, b - a);
    }
  };
  auto up = [&](void* dst, const void* src, size_t bytes) -> cudaError_t {
    if (bytes == 0) return cudaSuccess;
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->st);
  };
  CU(h, ensure(h->d_raw_xy, 2 * nl));
  CU(h, ensure(h->d_raw_pt, nl));
  CU(h, ensure(h->d_raw_a, nl));
  CU(h, ensure(h->d_raw_b, have_b ? nl : 1));
  CU(h, ensure(h->d_raw_in, nl));
  CU(h, ensure(h->d_bld_a, nl));
  CU(h, ensure(h->d_bld_flags, 8));
  // copy into the pinned arena and DMA in pieces, the largest array first: the link is busy from the first
  // 16 MB on, and only the last small piece is still in flight when the host cores are done
  auto stage_up = [&](void* dev, void* pinned, const void* src, size_t bytes) -> cudaError_t {
    constexpr size_t kPiece = size_t{16} << 20;
    for (size_t a = 0; a < bytes; a += kPiece) {
      const size_t len = std::min(kPiece, bytes - a);
      stage(static_cast<char*>(pinned) + a, static_cast<const char*>(src) + a, len);
      const cudaError_t e = up(static_cast<char*>(dev) + a, static_cast<char*>(pinned) + a, len);
      if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
  };
  CU(h, stage_up(h->d_raw_xy.p, s_xy, p->obs_xy + 2 * obs_lo, nl * 16));
  CU(h, stage_up(h->d_raw_pt.p, s_pt, p->obs_pt + obs_lo, nl * 4));
  CU(h, stage_up(h->d_raw_a.p, s_a, p->obs_pose_a + obs_lo, nl * 4));
  if (have_b) CU(h, stage_up(h->d_raw_b.p, s_b, p->obs_pose_b + obs_lo, nl * 4));
  CU(h, stage_up(h->d_raw_in.p, s_in, p->obs_intr + obs_lo, nl * 4));
  mark("stage + enqueue raw shard");
  // ---- per-observation records, validation, CSR offsets of the points
  h->n_ext = n_ext;
  h->n_intr = n_intr;
  CU(h, ensure(h->d_obs_xy, nl));
  CU(h, ensure(h->d_obs_ip, nl));
  CU(h, ensure(h->d_obs_ab, nl));
  CU(h, ensure(h->d_pt_first, n_pts + 1));
  CU(h, cudaMemsetAsync(h->d_bld_flags.p, 0, 8 * sizeof(int), h->st));
  const int prev_pt = obs_lo > 0 ? p->obs_pt[obs_lo - 1] : -1;
  bld_observations(nl, h->d_raw_xy.p, h->d_raw_pt.p, h->d_raw_a.p, have_b ? h->d_raw_b.p : nullptr, h->d_raw_in.p, pt_lo, n_pts, n_ext,
                   n_intr, prev_pt, h->d_obs_xy.p, h->d_obs_ip.p, h->d_obs_ab.p, h->d_bld_a.p, h->d_bld_flags.p, h->st);
  bld_lower_bound(h->d_raw_pt.p, nl, pt_lo, n_pts, h->d_pt_first.p, h->st);
  CU(h, cudaMemcpyAsync(s_first, h->d_pt_first.p, (static_cast<size_t>(n_pts) + 1) * sizeof(int), cudaMemcpyDeviceToHost, h->st));
  CU(h, cudaMemcpyAsync(s_flags, h->d_bld_flags.p, sizeof(int), cudaMemcpyDeviceToHost, h->st));
  CU(h, cudaStreamSynchronize(h->st));
  CU(h, cudaGetLastError());
  int flags = s_flags[0];
  // ranks take the same road: one of them meeting unsorted / composed / out-of-range input sends all to the host build
  double not_ok = (flags & (1 | 2 | 4)) ? 1.0 : 0.0;
  if (h->world > 1) {
    CU(h, ensure(h->d_scalars_red, S_TOTAL));
    CU(h, cudaMemcpyAsync(h->d_scalars_red.p, &not_ok, sizeof not_ok, cudaMemcpyHostToDevice, h->st));
    int rc = nccl_api().AllReduce(h->d_scalars_red.p, h->d_scalars_red.p, 1, kNcclFloat64, kNcclMax, h->comm, h->st);
    if (rc != 0) return h->fail(DBA_ERR_NCCL, "ncclAllReduce: %s", nccl_api().GetErrorString(rc));
    CU(h, cudaMemcpyAsync(&not_ok, h->d_scalars_red.p, sizeof not_ok, cudaMemcpyDeviceToHost, h->st));
    CU(h, cudaStreamSynchronize(h->st));
  }
  if (not_ok > 0.5) return DBA_OK;  // not handled here
  if (free_intr && (flags & 8)) return h->fail(DBA_ERR_UNSUPPORTED, "free_intrinsics needs obs_intr == obs_pose_a");
  mark("records + point offsets");
  // ---- tiles of whole points (host: O(#tiles) binary searches in the offsets), dense-solver batches
  const int* first = s_first;
  int max_track = 0;
#pragma omp parallel for schedule(static) reduction(max : max_track)
  for (int i = 0; i < n_pts; ++i) max_track = std::max(max_track, first[i + 1] - first[i]);
  if (h->world > 1) {  // the tile capacity follows the longest track of the WHOLE problem, as in the host build
    double mt = max_track;
    CU(h, cudaMemcpyAsync(h->d_scalars_red.p, &mt, sizeof mt, cudaMemcpyHostToDevice, h->st));
    int rc = nccl_api().AllReduce(h->d_scalars_red.p, h->d_scalars_red.p, 1, kNcclFloat64, kNcclMax, h->comm, h->st);
    if (rc != 0) return h->fail(DBA_ERR_NCCL, "ncclAllReduce: %s", nccl_api().GetErrorString(rc));
    CU(h, cudaMemcpyAsync(&mt, h->d_scalars_red.p, sizeof mt, cudaMemcpyDeviceToHost, h->st));
    CU(h, cudaStreamSynchronize(h->st));
    max_track = static_cast<int>(mt);
  }
  if (max_track > kMaxTile)
    return h->fail(DBA_ERR_UNSUPPORTED, "a point has %d observations; tracks longer than %d are not implemented", max_track, kMaxTile);
  int tile_cap = max_track <= 256 ? 256 : (max_track <= 512 ? 512 : 1024);
  if (tile_cap == 256 && !freeze && n / std::max(h->world, 1) >= int64_t{512} * 148 * 4) tile_cap = 512;
  if (const char* env = std::getenv("DBA_TILE")) {
    const int forced = std::atoi(env);
    if ((forced == 512 || forced == 1024) && forced >= tile_cap) tile_cap = forced;
    if (forced == 256 && max_track <= 256) tile_cap = 256;
  }
  std::vector<TileMeta> tile_meta;
  tile_meta.reserve(static_cast<size_t>(nl / 200 + 16));
  {
    const int max_pts = max_tile_points(tile_cap);
    for (int sp = 0; sp < n_pts;) {
      const int limit = first[sp] + tile_cap;
      int end = static_cast<int>(std::upper_bound(first + sp + 1, first + n_pts + 1, limit) - first) - 1;
      end = std::max(std::min(end, sp + max_pts), sp + 1);
      TileMeta m{};
      m.obs0 = first[sp];
      m.n_obs = first[end] - first[sp];
      m.pt0 = sp;
      m.n_pts = end - sp;
      tile_meta.push_back(m);
      sp = end;
    }
  }
  const int n_tiles = static_cast<int>(tile_meta.size());
  const bool dense_ok = cb > 0 && n_ext <= kDnMaxBlocks && n_ext * cb <= kDnMaxSize;
  std::vector<int> dn_batch;
  if (dense_ok) {
    int start = 0, ents = 0;
    dn_batch.push_back(0);
    for (int i = 0; i < n_pts; ++i) {
      const int c = std::min(2 * (first[i + 1] - first[i]), n_ext);
      if (i > start && (i - start >= kDnPtsCap || ents + c > kDnEntCap)) {
        dn_batch.push_back(i);
        start = i;
        ents = 0;
      }
      ents += c;
    }
    if (n_pts > 0) dn_batch.push_back(n_pts);
  }
  std::vector<int> tile_obs(static_cast<size_t>(n_tiles) + 1), tile_pt(static_cast<size_t>(n_tiles) + 1);
  for (int t = 0; t < n_tiles; ++t) {
    tile_obs[t] = tile_meta[t].obs0;
    tile_pt[t] = tile_meta[t].pt0;
  }
  tile_obs[n_tiles] = static_cast<int>(nl);
  tile_pt[n_tiles] = n_pts;
  mark("tiles");
  // ---- sizes known so far; the partial count comes from the device count pass
  h->freeze = freeze;
  h->free_intr = free_intr;
  h->cb = cb;
  h->two = 0;
  h->n_obs_global = n;
  h->n_pts_global = p->n_pts;
  h->pt_lo = pt_lo;
  h->n_pts = n_pts;
  h->n_obs = nl;
  BuildSizes z;
  z.nl = nl;
  z.ld = std::max<int64_t>(((nl + 63) / 64) * 64, 64);
  z.n_entries = cb ? nl : 0;
  z.n_pts = n_pts;
  z.n_tiles = n_tiles;
  z.tile_cap = tile_cap;
  z.n_cols = cb ? static_cast<size_t>(n_tiles) * tile_cap : 0;
  z.mf_w = cb == 9 ? 2 : 4;
  z.two = 0;
  z.cb = cb;
  z.intr_is_pose = (flags & 8) ? 0 : 1;
  z.dense_ok = dense_ok;
  z.n_dn_batches = dense_ok ? static_cast<int>(dn_batch.size()) - 1 : 0;
  CU(h, ensure(h->d_tile_meta, n_tiles));
  CU(h, ensure(h->d_tile_obs, n_tiles + 1));
  CU(h, ensure(h->d_tile_pt, n_tiles + 1));
  CU(h, cudaMemcpyAsync(h->d_tile_meta.p, tile_meta.data(), sizeof(TileMeta) * n_tiles, cudaMemcpyHostToDevice, h->st));
  CU(h, cudaMemcpyAsync(h->d_tile_obs.p, tile_obs.data(), sizeof(int) * (n_tiles + 1), cudaMemcpyHostToDevice, h->st));
  CU(h, cudaMemcpyAsync(h->d_tile_pt.p, tile_pt.data(), sizeof(int) * (n_tiles + 1), cudaMemcpyHostToDevice, h->st));
  const size_t temp_bytes = bld_temp_bytes(nl, static_cast<int>(std::min<int64_t>(nl, INT32_MAX)), n_tiles, n_ext);
  CU(h, ensure(h->d_bld_temp, temp_bytes));
  int n_partials = 0, n_chunks = 0;
  if (cb) {
    DeviceBuild B{};
    B.obs_ab = h->d_obs_ab.p;
    B.obs_ip = h->d_obs_ip.p;
    B.tile_meta = h->d_tile_meta.p;
    CU(h, ensure(h->d_bld_np, n_tiles + 1));
    CU(h, ensure(h->d_bld_g0, n_tiles + 1));
    CU(h, cudaMemsetAsync(h->d_bld_np.p, 0, sizeof(int) * (n_tiles + 1), h->st));
    B.tile_np = h->d_bld_np.p;
    bld_tiles(B, cb, tile_cap, n_tiles, /*fill=*/false, h->st);
    if (bld_exclusive_sum(h->d_bld_temp.p, temp_bytes, h->d_bld_np.p, h->d_bld_g0.p, n_tiles + 1, h->st) != 0)
      return h->fail(DBA_ERR_CUDA, "device scan failed");
    CU(h, cudaMemcpyAsync(s_flags + 1, h->d_bld_g0.p + n_tiles, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CU(h, cudaStreamSynchronize(h->st));
    n_partials = s_flags[1];
    z.n_partials = n_partials;
    // camera-sorted incidence first (its chunk count is the other size the buffers depend on)
    CU(h, ensure(h->d_cam_entries, nl));
    CU(h, ensure(h->d_bld_key, std::max<int64_t>(nl, n_partials)));
    CU(h, ensure(h->d_bld_id, std::max<int64_t>(nl, n_partials)));
    CU(h, ensure(h->d_bld_key2, std::max<int64_t>(nl, n_partials)));
    CU(h, ensure(h->d_bld_first, n_ext + 2));
    CU(h, ensure(h->d_bld_cnt, n_ext + 2));
    CU(h, ensure(h->d_cam_chunk_first, n_ext + 1));
    int key_bits = 1;
    while ((1 << key_bits) < n_ext && key_bits < 31) ++key_bits;
    bld_iota2(nl, h->d_bld_id.p, h->st);
    if (bld_sort_pairs(h->d_bld_temp.p, temp_bytes, h->d_bld_a.p, h->d_bld_key2.p, h->d_bld_id.p, h->d_cam_entries.p, static_cast<int>(nl),
                       key_bits, h->st) != 0)
      return h->fail(DBA_ERR_CUDA, "device sort failed");
    bld_lower_bound(h->d_bld_key2.p, nl, 0, n_ext, h->d_bld_first.p, h->st);
    CU(h, cudaMemsetAsync(h->d_bld_cnt.p, 0, sizeof(int) * (n_ext + 2), h->st));
    bld_chunk_counts(n_ext, h->d_bld_first.p, h->d_bld_cnt.p, h->st);
    if (bld_exclusive_sum(h->d_bld_temp.p, temp_bytes, h->d_bld_cnt.p, h->d_cam_chunk_first.p, n_ext + 1, h->st) != 0)
      return h->fail(DBA_ERR_CUDA, "device scan failed");
    CU(h, cudaMemcpyAsync(s_flags + 2, h->d_cam_chunk_first.p + n_ext, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CU(h, cudaStreamSynchronize(h->st));
    n_chunks = s_flags[2];
    z.n_chunks = n_chunks;
  }
  {
    const int rc = ensure_work_buffers(h, z);  // (buffers filled above are large enough already: grow only)
    if (rc != DBA_OK) return rc;
  }
  if (cb) {
    bld_chunks(n_ext, h->d_bld_first.p, h->d_cam_chunk_first.p, h->d_cam_chunks.p, h->st);
    // tile incidence: items, partial offsets, blocks, local indices, matrix-free columns
    DeviceBuild B{};
    B.obs_ab = h->d_obs_ab.p;
    B.obs_ip = h->d_obs_ip.p;
    B.tile_meta = h->d_tile_meta.p;
    B.tile_np = h->d_bld_np.p;
    B.tile_g0 = h->d_bld_g0.p;
    B.items = h->d_items.p;
    B.part_first_rel = h->d_part_first_rel.p;
    B.part_first = h->d_part_first.p;
    B.part_blk = h->d_part_blk.p;
    B.part_key = h->d_bld_key.p;
    B.part_id = h->d_bld_id.p;
    B.obs_lc = h->d_obs_lc.p;
    B.obs_lp = h->d_obs_lp.p;
    B.mf_cols = h->d_mf_cols.p;
    bld_tiles(B, cb, tile_cap, n_tiles, /*fill=*/true, h->st);
    // partials grouped by camera block: stable sort by block; row of partial g = its rank
    CU(h, ensure(h->d_bld_id2, std::max(n_partials, 1)));
    int key_bits = 1;
    while ((1 << key_bits) < n_ext && key_bits < 31) ++key_bits;
    if (n_partials > 0 && bld_sort_pairs(h->d_bld_temp.p, temp_bytes, h->d_bld_key.p, h->d_bld_key2.p, h->d_bld_id.p, h->d_bld_id2.p, n_partials,
                                         key_bits, h->st) != 0)
      return h->fail(DBA_ERR_CUDA, "device sort failed");
    bld_scatter_pos(n_partials, h->d_bld_id2.p, h->d_part_dst.p, h->st);
    bld_lower_bound(h->d_bld_key2.p, n_partials, 0, n_ext, h->d_cam_part_first.p, h->st);
  } else {
    // points only: the tile kernels still want the tile-local point of every observation
    DeviceBuild B{};
    B.obs_ip = h->d_obs_ip.p;
    B.tile_meta = h->d_tile_meta.p;
    B.obs_lc = h->d_obs_lc.p;
    B.obs_lp = h->d_obs_lp.p;
    bld_tiles_plain(B, n_tiles, h->st);
  }
  if (dense_ok) CU(h, cudaMemcpyAsync(h->d_dn_batch.p, dn_batch.data(), sizeof(int) * dn_batch.size(), cudaMemcpyHostToDevice, h->st));
  // points of the shard and the camera-side parameters
  stage(s_pts, p->pts + 3 * static_cast<size_t>(pt_lo), static_cast<size_t>(n_pts) * 24);
  CU(h, up(h->d_pts[2].p, s_pts, static_cast<size_t>(n_pts) * 24));
  {
    const int rc = upload_params(h, p);
    if (rc != DBA_OK) return rc;
  }
  h->perm.resize(static_cast<size_t>(nl));
  int64_t* perm = h->perm.data();
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < nl; ++k) perm[k] = obs_lo + k;
  CU(h, cudaStreamSynchronize(h->st));
  CU(h, cudaGetLastError());
  mark("device build + params");
  {
    const int rc = bind_problem(h, z);
    if (rc != DBA_OK) return rc;
  }
  // dba_problem_update reads the point-sorted observation image from the arena: for sorted single-pose input
  // that is the raw shard itself, re-packed lazily by dba_problem_update from the device copies
  h->keep.valid = false;
  h->build_mode = 1;
  h->have_problem = true;
  *handled = 1;
  return dba_params_reset(h);
}

// ------------------------------------------------------------------- problem upload
// Host side of the upload: every O(n_obs) step is a parallel pass (OpenMP over the host cores)
// writing straight into the pinned staging arena, copied with a handful of large async memcpys.
int dba_problem_set(dba_handle* h, const dba_problem* p) {
  if (!h || !p) return DBA_ERR_INVALID_ARGUMENT;
  CU(h, cudaSetDevice(h->device));
  h->have_problem = false;
  h->keep.valid = false;
  if (p->n_obs < 0 || p->n_pts < 0 || p->n_ext < 0 || p->n_intr < 0)
    return h->fail(DBA_ERR_INVALID_ARGUMENT, "negative size");
  if (p->n_obs > 0 && (!p->obs_xy || !p->obs_pt || !p->obs_pose_a || !p->obs_intr))
    return h->fail(DBA_ERR_INVALID_ARGUMENT, "null observation array");
  if ((p->n_pts > 0 && !p->pts) || (p->n_ext > 0 && (!p->ext_rot || !p->ext_trans)) ||
      (p->n_intr > 0 && (!p->intr_center || !p->intr_focal || !p->intr_dist || !p->intr_nf || !p->intr_nd)))
    return h->fail(DBA_ERR_INVALID_ARGUMENT, "null parameter array");
  if (p->n_obs >= (int64_t{1} << 30)) return h->fail(DBA_ERR_UNSUPPORTED, "more than 2^30 observations per handle");
  const int64_t n = p->n_obs;
  const int n_ext = p->n_ext, n_intr = p->n_intr;
  h->build_mode = 0;
  {
    // device-side construction, the default: every rank stages and builds only its shard, and a
    // single GPU builds a 5M-observation problem in half the time of the 16-core host build.  Point-sorted
    // single-pose input only; otherwise (or when any rank says so) the host build below runs.
    // DBA_BUILD=host forces the host build.
    const char* env = std::getenv("DBA_BUILD");
    const bool want_device = env ? std::strcmp(env, "host") != 0 : true;
    if (want_device) {
      int handled = 0;
      const int rc = problem_set_device(h, p, &handled);
      if (rc != DBA_OK || handled) return rc;
    }
  }

  OmpThreadScope omp_scope(h->world);
  const bool timing = std::getenv("DBA_TIMING") != nullptr;
  double t_mark = now_s();
  auto mark = [&](const char* what) {
    if (!timing) return;
    const double t = now_s();
    std::fprintf(stderr, "[dba_problem_set] %-28s %8.2f ms\n", what, 1e3 * (t - t_mark));
    t_mark = t;
  };
  // ---- pass 1 (parallel): validation, two-pose detection, sortedness
  int64_t bad_index = -1;
  int two = 0, sorted = 1, intr_is_pose = 1;
#pragma omp parallel for schedule(static) reduction(max : bad_index, two) reduction(min : sorted, intr_is_pose)
  for (int64_t i = 0; i < n; ++i) {
    const int pt = p->obs_pt[i], a = p->obs_pose_a[i], in = p->obs_intr[i];
    const int b = p->obs_pose_b ? p->obs_pose_b[i] : -1;
    if (pt < 0 || pt >= p->n_pts || a < 0 || a >= n_ext || in < 0 || in >= n_intr || b < -1 || b >= n_ext) bad_index = std::max(bad_index, i);
    if (b >= 0) two = 1;
    if (i > 0 && p->obs_pt[i - 1] > pt) sorted = 0;
    if (in != a) intr_is_pose = 0;
  }
  if (bad_index >= 0) return h->fail(DBA_ERR_INVALID_ARGUMENT, "observation %lld: index out of range", (long long)bad_index);
  for (int i = 0; i < n_intr; ++i)
    if (p->intr_nf[i] < 1 || p->intr_nf[i] > 2 || p->intr_nd[i] < 0 || p->intr_nd[i] > 2)
      return h->fail(DBA_ERR_INVALID_ARGUMENT, "intrinsic %d: nf must be 1|2 and nd 0|1|2", i);
  h->freeze = p->freeze_camera != 0;
  h->free_intr = (!h->freeze && p->free_intrinsics) ? 1 : 0;
  if (h->free_intr) {
    if (two) return h->fail(DBA_ERR_UNSUPPORTED, "free_intrinsics with composed poses is not implemented");
    if (n_ext != n_intr) return h->fail(DBA_ERR_UNSUPPORTED, "free_intrinsics needs one intrinsic per extrinsic");
    for (int i = 0; i < n_intr; ++i)
      if (p->intr_nf[i] != 1 || p->intr_nd[i] != 2) return h->fail(DBA_ERR_UNSUPPORTED, "free_intrinsics needs nf=1, nd=2");
    if (!intr_is_pose) return h->fail(DBA_ERR_UNSUPPORTED, "free_intrinsics needs obs_intr == obs_pose_a");
  }
  h->cb = h->freeze ? 0 : (h->free_intr ? 9 : 6);
  h->two = two;
  h->n_obs_global = n;
  h->n_pts_global = p->n_pts;
  h->n_ext = n_ext;
  h->n_intr = n_intr;
  const int cb = h->cb;

  mark("validate");
  // ---- observations per point (global), shard of this rank
  std::vector<int64_t> pt_count(static_cast<size_t>(p->n_pts) + 1, 0);
  if (sorted) {
    // run boundaries of a sorted key array: first[pt] = first position with key >= pt
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i <= n; ++i) {
      const int lo = i == 0 ? 0 : p->obs_pt[i - 1] + 1;
      const int hi = i == n ? p->n_pts : p->obs_pt[i];
      for (int q = lo; q <= hi; ++q) pt_count[q] = i;
    }
  } else {
    // unsorted input (the usual case for a scene graph in file order): stable counting sort by point,
    // lock-free — every thread scans all observations and counts those of ITS range of points
#pragma omp parallel
    {
      const int t = omp_get_thread_num(), nt = omp_get_num_threads();
      const int p0 = static_cast<int>(static_cast<int64_t>(p->n_pts) * t / nt);
      const int p1 = static_cast<int>(static_cast<int64_t>(p->n_pts) * (t + 1) / nt);
      for (int64_t i = 0; i < n; ++i) {
        const int pt = p->obs_pt[i];
        if (pt >= p0 && pt < p1) pt_count[pt + 1]++;
      }
    }
    for (int i = 0; i < p->n_pts; ++i) pt_count[i + 1] += pt_count[i];
  }
  int pt_lo = 0, pt_hi = p->n_pts;
  if (h->world > 1) {
    // same rule as dba_shard_plan
    auto cut = [&](int r) -> int {
      if (r <= 0) return 0;
      if (r >= h->world) return p->n_pts;
      const int64_t target = n * r / h->world;
      return std::min(static_cast<int>(std::lower_bound(pt_count.begin(), pt_count.end(), target) - pt_count.begin()), p->n_pts);
    };
    pt_lo = cut(h->rank);
    for (int r = 1; r < h->rank; ++r) pt_lo = std::max(pt_lo, cut(r));
    pt_hi = std::max(cut(h->rank + 1), pt_lo);
    if (h->rank + 1 == h->world) pt_hi = p->n_pts;
  }
  h->pt_lo = pt_lo;
  h->n_pts = pt_hi - pt_lo;
  const int n_pts = h->n_pts;
  const int64_t obs_lo = pt_count[pt_lo], obs_hi = pt_count[pt_hi];
  const int64_t nl = obs_hi - obs_lo;
  h->n_obs = nl;

  // ---- permutation (local sorted position -> caller's index)
  h->perm.resize(static_cast<size_t>(nl));
  if (sorted) {
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < nl; ++k) h->perm[k] = obs_lo + k;
  } else {
    // second pass of the counting sort, same split: a thread places the observations of its points in
    // index order (stable), so no two threads touch the same cursor
    std::vector<int64_t> cursor(pt_count.begin() + pt_lo, pt_count.begin() + pt_hi);
#pragma omp parallel
    {
      const int t = omp_get_thread_num(), nt = omp_get_num_threads();
      const int q0 = pt_lo + static_cast<int>(static_cast<int64_t>(pt_hi - pt_lo) * t / nt);
      const int q1 = pt_lo + static_cast<int>(static_cast<int64_t>(pt_hi - pt_lo) * (t + 1) / nt);
      for (int64_t i = 0; i < n; ++i) {
        const int pt = p->obs_pt[i];
        if (pt >= q0 && pt < q1) h->perm[cursor[pt - pt_lo]++ - obs_lo] = i;
      }
    }
  }
  const int64_t* perm = h->perm.data();

  mark("point counts + permutation");
  // ---- tiles of whole points (greedy, serial: O(n_pts))
  // tile capacity: 256 unless some point has a longer track (then 512 / 1024; two-pose
  // problems stop at 512 because a 1024-observation two-pose tile exceeds shared memory)
  int64_t max_track = 0;
#pragma omp parallel for schedule(static) reduction(max : max_track)
  for (int i = 0; i < p->n_pts; ++i) max_track = std::max(max_track, pt_count[i + 1] - pt_count[i]);
  const int tile_cap_max = (two && !h->freeze) ? 512 : kMaxTile;
  if (max_track > tile_cap_max)
    return h->fail(DBA_ERR_UNSUPPORTED, "a point has %lld observations; tracks longer than %d are not implemented",
                   (long long)max_track, tile_cap_max);
  int tile_cap = max_track <= 256 ? 256 : (max_track <= 512 ? 512 : 1024);
  // single-pose problems large enough to fill the machine twice over use 512-observation tiles:
  // half as many (tile, camera) partials and camera-row fetches per observation (measured on bal5m)
  if (tile_cap == 256 && !two && !h->freeze && n / std::max(h->world, 1) >= int64_t{512} * 148 * 4) tile_cap = 512;
  if (const char* env = std::getenv("DBA_TILE")) {  // tuning knob: force a larger tile capacity
    const int forced = std::atoi(env);
    if ((forced == 512 || forced == 1024) && forced >= tile_cap && forced <= tile_cap_max) tile_cap = forced;
    if (forced == 256 && max_track <= 256) tile_cap = 256;
  }
  std::vector<TileMeta> tile_meta;
  tile_meta.reserve(static_cast<size_t>(nl / 200 + 16));
  {
    // greedy packing = from each tile start, the longest run of whole points within the capacity:
    // one binary search in the prefix sums per TILE instead of one step per point
    const int64_t* first = pt_count.data() + pt_lo;  // first[i] = first observation of local point i (global position)
    const int max_pts = max_tile_points(tile_cap);
    for (int s = 0; s < n_pts;) {
      const int64_t limit = first[s] + tile_cap;
      int end = static_cast<int>(std::upper_bound(first + s + 1, first + n_pts + 1, limit) - first) - 1;
      end = std::max(std::min(end, s + max_pts), s + 1);
      TileMeta m{};
      m.obs0 = static_cast<int>(first[s] - obs_lo);
      m.n_obs = static_cast<int>(first[end] - first[s]);
      m.pt0 = s;
      m.n_pts = end - s;
      tile_meta.push_back(m);
      s = end;
    }
  }
  const int n_tiles = static_cast<int>(tile_meta.size());

  mark("tiles");
  // ---- sizes of everything that goes through the pinned arena
  const int64_t ld = std::max<int64_t>(((nl + 63) / 64) * 64, 64);
  const int64_t n_ent_max = cb ? 2 * nl : 0;
  size_t arena_bytes = 0;
  auto want = [&](size_t bytes) { arena_bytes += pad256(bytes) + 256; };
  want(nl * sizeof(double2));          // obs_xy
  want(nl * sizeof(int2));             // obs_ip
  want(nl * sizeof(int2));             // obs_ab
  want(nl * sizeof(unsigned short));   // obs_lp
  want((n_pts + 1) * sizeof(int));     // pt_first
  want(n_tiles * sizeof(TileMeta));
  want((n_tiles + 1) * sizeof(int) * 3);
  want(n_ent_max * sizeof(int));       // cam_entries
  want(n_ent_max * sizeof(unsigned short));  // items
  want((n_ent_max + n_tiles + 1) * sizeof(unsigned short));  // part_first_rel
  want(n_ent_max * sizeof(int));        // part_dst
  want(n_ent_max * sizeof(int));        // part_blk
  want(nl * sizeof(ushort2));           // obs_lc
  want(cb ? (static_cast<size_t>(n_tiles) * tile_cap) * sizeof(int4) : 0);  // mf_cols (padded per tile)
  want((two && cb) ? n_ent_max * sizeof(int) : 0);  // items_mf
  want((n_ent_max + n_tiles + 1) * sizeof(int));    // part_first
  want((n_ent_max / 1024 + n_ext + 2) * sizeof(int4));  // cam_chunks
  want((n_ext + 1) * sizeof(int) * 3);
  // dense reduced system (DENSE_SCHUR): eligible when the camera side is small
  const bool dense_ok = cb > 0 && n_ext <= kDnMaxBlocks && n_ext * cb <= kDnMaxSize;
  want(dense_ok ? (static_cast<size_t>(n_pts) + 2) * sizeof(int) : 0);  // dn_batch
  const size_t dn_pairs = dense_ok ? static_cast<size_t>(n_ext) * (n_ext + 1) / 2 : 0;
  want((dense_ok && two) ? nl * sizeof(int) : 0);                             // dn_pair_entries
  want((dense_ok && two) ? (nl / 1024 + dn_pairs + 2) * sizeof(int4) : 0);    // dn_pair_chunks
  want((dense_ok && two) ? (dn_pairs + 2) * sizeof(int) : 0);                 // dn_pair_chunk_first
  want(3 * sizeof(double) * static_cast<size_t>(h->n_pts));  // the caller's points (pageable) staged for the async copy
  PinnedArena& A = arena_of(h);
  CU(h, cudaStreamSynchronize(h->st));  // the arena may still feed copies of a previous call
  CU(h, A.reserve(arena_bytes));

  mark("arena reserve");
  // Copies are enqueued as soon as a group of arrays is complete, so the DMA engine works while
  // the host builds the next group (everything staged is pinned: the copies are truly asynchronous).
  auto up = [&](void* dst, const void* src, size_t bytes) -> cudaError_t {
    if (bytes == 0) return cudaSuccess;
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->st);
  };
  // ---- per-observation arrays (parallel)
  double2* s_xy = A.take<double2>(nl);
  int2* s_ip = A.take<int2>(nl);
  int2* s_ab = A.take<int2>(nl);
  unsigned short* s_lp = A.take<unsigned short>(nl);
  int* s_pt_first = A.take<int>(static_cast<size_t>(n_pts) + 1);
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < nl; ++k) {
    const int64_t i = perm[k];
    s_xy[k] = make_double2(p->obs_xy[2 * i], p->obs_xy[2 * i + 1]);
    s_ip[k] = make_int2(p->obs_intr[i], p->obs_pt[i] - pt_lo);
    s_ab[k] = make_int2(p->obs_pose_a[i], p->obs_pose_b ? p->obs_pose_b[i] : -1);
  }
#pragma omp parallel for schedule(static)
  for (int i = 0; i <= n_pts; ++i) s_pt_first[i] = static_cast<int>(pt_count[pt_lo + i] - obs_lo);

  CU(h, ensure(h->d_obs_xy, nl));
  CU(h, ensure(h->d_obs_ip, nl));
  CU(h, ensure(h->d_obs_ab, nl));
  CU(h, ensure(h->d_pt_first, n_pts + 1));
  CU(h, up(h->d_obs_xy.p, s_xy, nl * sizeof(double2)));
  CU(h, up(h->d_obs_ip.p, s_ip, nl * sizeof(int2)));
  CU(h, up(h->d_obs_ab.p, s_ab, nl * sizeof(int2)));
  CU(h, up(h->d_pt_first.p, s_pt_first, (static_cast<size_t>(n_pts) + 1) * sizeof(int)));
  {
    // the caller's points: pageable -> pinned with all host cores, then one asynchronous copy
    // (slot 2 = pristine copy for dba_params_reset)
    const size_t n3 = 3 * static_cast<size_t>(n_pts);
    for (int sl = 0; sl < 3; ++sl) CU(h, ensure(h->d_pts[sl], n3));
    double* s_pts = A.take<double>(std::max<size_t>(n3, 1));
    const double* src = p->pts + 3 * static_cast<size_t>(pt_lo);
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < static_cast<int64_t>((n3 + 65535) / 65536); ++c) {
      const size_t a = static_cast<size_t>(c) * 65536, b = std::min(n3, a + 65536);
      std::memcpy(s_pts + a, src + a, (b - a) * sizeof(double));
    }
    CU(h, up(h->d_pts[2].p, s_pts, n3 * sizeof(double)));
  }
  mark("per-observation arrays");
  // ---- camera-sorted incidence: stable parallel counting sort of (obs, slot) entries by block
  int* s_cam_entries = A.take<int>(static_cast<size_t>(n_ent_max));
  int4* s_cam_chunks = nullptr;
  int* s_cam_chunk_first = A.take<int>(static_cast<size_t>(n_ext) + 1);
  int64_t n_entries = 0;
  int n_chunks = 0;
  const int kChunk = 1024;
  if (cb) {
    const int nthreads = std::max(1, omp_get_max_threads());
    std::vector<std::vector<int>> hist(nthreads, std::vector<int>(static_cast<size_t>(n_ext) + 1, 0));
#pragma omp parallel num_threads(nthreads)
    {
      const int t = omp_get_thread_num();
      const int64_t k0 = nl * t / nthreads, k1 = nl * (t + 1) / nthreads;
      std::vector<int>& hc = hist[t];
      for (int64_t k = k0; k < k1; ++k) {
        hc[s_ab[k].x]++;
        if (s_ab[k].y >= 0) hc[s_ab[k].y]++;
      }
    }
    // offsets: for block b, thread t starts at first[b] + sum_{t' < t} hist[t'][b]
    std::vector<int64_t> first(static_cast<size_t>(n_ext) + 1, 0);
    for (int b = 0; b < n_ext; ++b) {
      int64_t tot = 0;
      for (int t = 0; t < nthreads; ++t) {
        const int c = hist[t][b];
        hist[t][b] = static_cast<int>(tot);
        tot += c;
      }
      first[b + 1] = first[b] + tot;
    }
    n_entries = first[n_ext];
#pragma omp parallel num_threads(nthreads)
    {
      const int t = omp_get_thread_num();
      const int64_t k0 = nl * t / nthreads, k1 = nl * (t + 1) / nthreads;
      std::vector<int>& hc = hist[t];
      for (int64_t k = k0; k < k1; ++k) {
        const int a = s_ab[k].x, b = s_ab[k].y;
        const int64_t pa = first[a] + hc[a]++;
        s_cam_entries[pa] = static_cast<int>(2 * k);
        if (b >= 0) {
          const int64_t pb = first[b] + hc[b]++;
          s_cam_entries[pb] = static_cast<int>(2 * k + 1);
        }
      }
    }
    for (int b = 0; b < n_ext; ++b) n_chunks += static_cast<int>((first[b + 1] - first[b] + kChunk - 1) / kChunk);
    s_cam_chunks = A.take<int4>(static_cast<size_t>(std::max(n_chunks, 1)));
    int c = 0;
    for (int b = 0; b < n_ext; ++b) {
      s_cam_chunk_first[b] = c;
      for (int64_t e = first[b]; e < first[b + 1]; e += kChunk)
        s_cam_chunks[c++] = make_int4(b, static_cast<int>(e), static_cast<int>(std::min<int64_t>(e + kChunk, first[b + 1])), 0);
    }
    s_cam_chunk_first[n_ext] = c;
  } else {
    for (int b = 0; b <= n_ext; ++b) s_cam_chunk_first[b] = 0;
    s_cam_chunks = A.take<int4>(1);
  }

  CU(h, ensure(h->d_cam_entries, static_cast<size_t>(n_entries)));
  CU(h, ensure(h->d_cam_chunks, n_chunks));
  CU(h, ensure(h->d_cam_chunk_first, n_ext + 1));
  CU(h, up(h->d_cam_entries.p, s_cam_entries, n_entries * sizeof(int)));
  CU(h, up(h->d_cam_chunks.p, s_cam_chunks, n_chunks * sizeof(int4)));
  CU(h, up(h->d_cam_chunk_first.p, s_cam_chunk_first, (n_ext + 1) * sizeof(int)));
  mark("camera-sorted incidence");
  // ---- static tile-local camera incidence (parallel over tiles, two passes)
  unsigned short* s_items = A.take<unsigned short>(static_cast<size_t>(std::max<int64_t>(n_entries, 1)));
  unsigned short* s_part_first_rel = A.take<unsigned short>(static_cast<size_t>(n_entries + n_tiles + 1));
  int* s_cam_part_first = A.take<int>(static_cast<size_t>(n_ext) + 1);
  int* s_part_dst = nullptr;
  const int mf_w = cb == 9 ? 2 : 4;  // ints per column record
  const size_t n_cols = cb ? static_cast<size_t>(n_tiles) * tile_cap : 0;  // padded: tile t owns [t * tile_cap, (t + 1) * tile_cap)
  int* s_mf_cols = A.take<int>(std::max<size_t>(n_cols * mf_w, 1));
  int* s_items_mf = A.take<int>((two && cb) ? static_cast<size_t>(std::max<int64_t>(n_entries, 1)) : 1);
  int* s_part_first = A.take<int>(static_cast<size_t>(n_entries + n_tiles + 1));
  ushort2* s_obs_lc = A.take<ushort2>(static_cast<size_t>(std::max<int64_t>(nl, 1)));
  int n_partials = 0;
  std::vector<int> part_block;
  {
    // pass A: distinct blocks and items per tile
    std::vector<int> tile_np(static_cast<size_t>(n_tiles) + 1, 0), tile_ni(static_cast<size_t>(n_tiles) + 1, 0);
    if (cb) {
#pragma omp parallel
      {
        std::vector<int> seen(static_cast<size_t>(n_ext), -1);
#pragma omp for schedule(static)
        for (int t = 0; t < n_tiles; ++t) {
          const TileMeta& m = tile_meta[t];
          int np = 0, ni = 0;
          for (int k = m.obs0; k < m.obs0 + m.n_obs; ++k) {
            const int a = s_ab[k].x, b = s_ab[k].y;
            if (seen[a] != t) { seen[a] = t; ++np; }
            ++ni;
            if (b >= 0) {
              if (seen[b] != t) { seen[b] = t; ++np; }
              ++ni;
            }
          }
          tile_np[t + 1] = np;
          tile_ni[t + 1] = ni;
        }
      }
    }
    for (int t = 0; t < n_tiles; ++t) {
      tile_np[t + 1] += tile_np[t];
      tile_ni[t + 1] += tile_ni[t];
    }
    n_partials = tile_np[n_tiles];
    part_block.resize(static_cast<size_t>(n_partials));
    // pass B: fill (items grouped by local block, in order of first appearance)
#pragma omp parallel
    {
      std::vector<int> local_of(static_cast<size_t>(n_ext), -1), locals, count, start, col_of;
#pragma omp for schedule(static)
      for (int t = 0; t < n_tiles; ++t) {
        TileMeta& m = tile_meta[t];
        m.g0 = tile_np[t];
        m.n_parts = tile_np[t + 1] - tile_np[t];
        m.item0 = tile_ni[t];
        m.n_items = tile_ni[t + 1] - tile_ni[t];
        for (int k = m.obs0; k < m.obs0 + m.n_obs; ++k) {
          s_lp[k] = static_cast<unsigned short>(s_ip[k].y - m.pt0);
          s_obs_lc[k] = make_ushort2(0xffff, 0xffff);
        }
        if (!cb) continue;
        locals.clear();
        count.clear();
        for (int k = m.obs0; k < m.obs0 + m.n_obs; ++k) {
          for (int slot = 0; slot < 2; ++slot) {
            const int blk = slot ? s_ab[k].y : s_ab[k].x;
            if (blk < 0) continue;
            int lc = local_of[blk];
            if (lc < 0) {
              lc = static_cast<int>(locals.size());
              local_of[blk] = lc;
              locals.push_back(blk);
              count.push_back(0);
            }
            count[lc]++;
          }
        }
        start.assign(locals.size() + 1, 0);
        for (size_t lc = 0; lc < locals.size(); ++lc) start[lc + 1] = start[lc] + count[lc];
        unsigned short* rel = s_part_first_rel + m.g0 + t;
        for (size_t lc = 0; lc <= locals.size(); ++lc) {
          rel[lc] = static_cast<unsigned short>(start[lc]);
          s_part_first[m.g0 + t + lc] = start[lc];
        }
        for (size_t lc = 0; lc < locals.size(); ++lc) part_block[m.g0 + lc] = locals[lc];
        for (int k = m.obs0; k < m.obs0 + m.n_obs; ++k) {
          s_obs_lc[k] = make_ushort2(static_cast<unsigned short>(local_of[s_ab[k].x]),
                                     s_ab[k].y >= 0 ? static_cast<unsigned short>(local_of[s_ab[k].y]) : 0xffff);
          const int lo = k - m.obs0;
          for (int slot = 0; slot < 2; ++slot) {
            const int blk = slot ? s_ab[k].y : s_ab[k].x;
            if (blk < 0) continue;
            const int lc = local_of[blk];
            s_items[m.item0 + start[lc]++] = static_cast<unsigned short>(lo | (slot << 15));
          }
        }
        for (int blk : locals) local_of[blk] = -1;
        // columns of the matrix-free product: the tile's observations in camera-block order
        // (slot-0 items in item order; every observation has exactly one)
        col_of.assign(static_cast<size_t>(m.n_obs), 0);
        int col = 0;
        for (int i = 0; i < m.n_items; ++i) {
          const unsigned short it = s_items[m.item0 + i];
          if (it >> 15) continue;
          const int lo = it & 0x7fff, k = m.obs0 + lo;
          col_of[lo] = col;
          int* rec = s_mf_cols + (static_cast<size_t>(t) * tile_cap + col) * mf_w;
          const int lplo = static_cast<int>((static_cast<unsigned int>(s_lp[k]) << 16) | static_cast<unsigned int>(lo));
          rec[0] = s_ab[k].x;
          if (mf_w == 2) {
            rec[1] = lplo;
          } else {
            rec[1] = s_ab[k].y;
            rec[2] = lplo;
            rec[3] = s_ip[k].x;
          }
          ++col;
        }
        for (; col < tile_cap; ++col) {
          int* rec = s_mf_cols + (static_cast<size_t>(t) * tile_cap + col) * mf_w;
          for (int w = 0; w < mf_w; ++w) rec[w] = w == 0 ? -1 : 0;
        }
        if (two)
          for (int i = 0; i < m.n_items; ++i) {
            const unsigned short it = s_items[m.item0 + i];
            s_items_mf[m.item0 + i] = col_of[it & 0x7fff] | (it & 0x8000);
          }
      }
    }
    // partials grouped by camera block (serial counting sort: O(n_partials))
    for (int b = 0; b <= n_ext; ++b) s_cam_part_first[b] = 0;
    for (int g = 0; g < n_partials; ++g) s_cam_part_first[part_block[g] + 1]++;
    for (int b = 0; b < n_ext; ++b) s_cam_part_first[b + 1] += s_cam_part_first[b];
    std::vector<int> cur(s_cam_part_first, s_cam_part_first + n_ext);
    s_part_dst = A.take<int>(static_cast<size_t>(std::max(n_partials, 1)));
    for (int g = 0; g < n_partials; ++g) {
      const int pos = cur[part_block[g]]++;
      s_part_dst[g] = pos;
    }
  }
  TileMeta* s_tile_meta = A.take<TileMeta>(static_cast<size_t>(std::max(n_tiles, 1)));
  if (n_tiles) std::memcpy(s_tile_meta, tile_meta.data(), sizeof(TileMeta) * n_tiles);
  int* s_tile_obs = A.take<int>(static_cast<size_t>(n_tiles) + 1);
  int* s_tile_pt = A.take<int>(static_cast<size_t>(n_tiles) + 1);
  for (int t = 0; t < n_tiles; ++t) {
    s_tile_obs[t] = tile_meta[t].obs0;
    s_tile_pt[t] = tile_meta[t].pt0;
  }
  s_tile_obs[n_tiles] = static_cast<int>(nl);
  s_tile_pt[n_tiles] = n_pts;

  // ---- point batches of the dense reduced-system kernel: greedy runs of whole points with at most
  // kDnPtsCap points and kDnEntCap (point, camera block) entries, entry capacity min(2 k_i, n_blocks)
  int n_dn_batches = 0;
  int* s_dn_batch = nullptr;
  if (dense_ok) {
    s_dn_batch = A.take<int>(static_cast<size_t>(n_pts) + 2);
    const int64_t* first = pt_count.data() + pt_lo;
    int start = 0, ents = 0;
    s_dn_batch[0] = 0;
    for (int i = 0; i < n_pts; ++i) {
      const int c = static_cast<int>(std::min<int64_t>(2 * (first[i + 1] - first[i]), n_ext));
      if (i > start && (i - start >= kDnPtsCap || ents + c > kDnEntCap)) {
        s_dn_batch[++n_dn_batches] = i;
        start = i;
        ents = 0;
      }
      ents += c;
    }
    if (n_pts > 0) s_dn_batch[++n_dn_batches] = n_pts;
  }
  // composed observations sorted by their (lower, higher) pose-block pair, in chunks of <= 1024:
  // k_pair_gather sums F_lo^T F_hi per chunk (stable parallel counting sort, as for the camera incidence)
  int n_dn_pair_chunks = 0;
  int64_t n_dn_pair_entries = 0;
  int *s_dn_pair_entries = nullptr, *s_dn_pair_chunk_first = nullptr;
  int4* s_dn_pair_chunks = nullptr;
  if (dense_ok && two) {
    const int np = static_cast<int>(dn_pairs);
    auto pair_of = [n_ext](int a, int b) { return a * n_ext - a * (a - 1) / 2 + (b - a); };  // a <= b
    const int nthreads = std::max(1, omp_get_max_threads());
    std::vector<std::vector<int>> hist(nthreads, std::vector<int>(static_cast<size_t>(np) + 1, 0));
#pragma omp parallel num_threads(nthreads)
    {
      const int t = omp_get_thread_num();
      const int64_t k0 = nl * t / nthreads, k1 = nl * (t + 1) / nthreads;
      for (int64_t k = k0; k < k1; ++k) {
        const int a = s_ab[k].x, b = s_ab[k].y;
        if (b >= 0 && b != a) hist[t][pair_of(std::min(a, b), std::max(a, b))]++;
      }
    }
    std::vector<int64_t> first(static_cast<size_t>(np) + 1, 0);
    for (int q = 0; q < np; ++q) {
      int64_t tot = 0;
      for (int t = 0; t < nthreads; ++t) {
        const int c = hist[t][q];
        hist[t][q] = static_cast<int>(tot);
        tot += c;
      }
      first[q + 1] = first[q] + tot;
    }
    n_dn_pair_entries = first[np];
    s_dn_pair_entries = A.take<int>(static_cast<size_t>(std::max<int64_t>(n_dn_pair_entries, 1)));
#pragma omp parallel num_threads(nthreads)
    {
      const int t = omp_get_thread_num();
      const int64_t k0 = nl * t / nthreads, k1 = nl * (t + 1) / nthreads;
      for (int64_t k = k0; k < k1; ++k) {
        const int a = s_ab[k].x, b = s_ab[k].y;
        if (b < 0 || b == a) continue;
        const int q = pair_of(std::min(a, b), std::max(a, b));
        s_dn_pair_entries[first[q] + hist[t][q]++] = static_cast<int>(2 * k + (b < a ? 1 : 0));
      }
    }
    for (int q = 0; q < np; ++q) n_dn_pair_chunks += static_cast<int>((first[q + 1] - first[q] + 1023) / 1024);
    s_dn_pair_chunks = A.take<int4>(static_cast<size_t>(std::max(n_dn_pair_chunks, 1)));
    s_dn_pair_chunk_first = A.take<int>(static_cast<size_t>(np) + 1);
    int c = 0;
    for (int q = 0; q < np; ++q) {
      s_dn_pair_chunk_first[q] = c;
      for (int64_t e = first[q]; e < first[q + 1]; e += 1024)
        s_dn_pair_chunks[c++] = make_int4(q, static_cast<int>(e), static_cast<int>(std::min<int64_t>(e + 1024, first[q + 1])), 0);
    }
    s_dn_pair_chunk_first[np] = c;
  }
  mark("tile incidence + columns");
  // ---- device buffers (kept across calls, grow only), host-built arrays -> device, binding
  BuildSizes z;
  z.nl = nl;
  z.ld = ld;
  z.n_entries = n_entries;
  z.n_pts = n_pts;
  z.n_tiles = n_tiles;
  z.tile_cap = tile_cap;
  z.n_chunks = n_chunks;
  z.n_partials = n_partials;
  z.n_cols = n_cols;
  z.mf_w = mf_w;
  z.two = two;
  z.cb = cb;
  z.intr_is_pose = intr_is_pose;
  z.dense_ok = dense_ok;
  z.n_dn_batches = n_dn_batches;
  {
    const int rc = ensure_work_buffers(h, z);
    if (rc != DBA_OK) return rc;
  }
  if (dense_ok) {
    CU(h, up(h->d_dn_batch.p, s_dn_batch, (static_cast<size_t>(n_dn_batches) + 1) * sizeof(int)));
    if (two && n_dn_pair_chunks > 0) {
      CU(h, ensure(h->d_dn_pair_entries, static_cast<size_t>(n_dn_pair_entries)));
      CU(h, ensure(h->d_dn_pair_chunks, static_cast<size_t>(n_dn_pair_chunks)));
      CU(h, ensure(h->d_dn_pair_chunk_first, dn_pairs + 1));
      CU(h, ensure(h->d_dn_pair_acc, static_cast<size_t>(n_dn_pair_chunks) * 36));
      CU(h, up(h->d_dn_pair_entries.p, s_dn_pair_entries, n_dn_pair_entries * sizeof(int)));
      CU(h, up(h->d_dn_pair_chunks.p, s_dn_pair_chunks, n_dn_pair_chunks * sizeof(int4)));
      CU(h, up(h->d_dn_pair_chunk_first.p, s_dn_pair_chunk_first, (dn_pairs + 1) * sizeof(int)));
      h->Q.pair_entries = h->d_dn_pair_entries.p;
      h->Q.pair_chunks = h->d_dn_pair_chunks.p;
      h->Q.pair_chunk_first = h->d_dn_pair_chunk_first.p;
      h->Q.pair_chunk_acc = h->d_dn_pair_acc.p;
      h->Q.n_pair_chunks = n_dn_pair_chunks;
    }
  }
  CU(h, up(h->d_obs_lp.p, s_lp, nl * sizeof(unsigned short)));
  CU(h, up(h->d_tile_obs.p, s_tile_obs, (n_tiles + 1) * sizeof(int)));
  CU(h, up(h->d_tile_pt.p, s_tile_pt, (n_tiles + 1) * sizeof(int)));
  CU(h, up(h->d_tile_meta.p, s_tile_meta, n_tiles * sizeof(TileMeta)));
  CU(h, up(h->d_part_first_rel.p, s_part_first_rel, (n_entries + n_tiles + 1) * sizeof(unsigned short)));
  CU(h, up(h->d_items.p, s_items, n_entries * sizeof(unsigned short)));
  CU(h, up(h->d_cam_part_first.p, s_cam_part_first, (n_ext + 1) * sizeof(int)));
  CU(h, up(h->d_part_dst.p, s_part_dst, n_partials * sizeof(int)));
  {
    int* s_part_blk = A.take<int>(static_cast<size_t>(std::max(n_partials, 1)));
    if (n_partials) std::memcpy(s_part_blk, part_block.data(), sizeof(int) * n_partials);
    CU(h, up(h->d_part_blk.p, s_part_blk, n_partials * sizeof(int)));
    CU(h, up(h->d_obs_lc.p, s_obs_lc, nl * sizeof(ushort2)));
  }
  if (cb) CU(h, up(h->d_mf_cols.p, s_mf_cols, n_cols * mf_w * sizeof(int)));
  if (two && cb) CU(h, up(h->d_items_mf.p, s_items_mf, n_entries * sizeof(int)));
  if (cb) CU(h, up(h->d_part_first.p, s_part_first, (n_entries + n_tiles + 1) * sizeof(int)));
  {
    const int rc = upload_params(h, p);
    if (rc != DBA_OK) return rc;
  }
  mark("alloc + enqueue copies");
  CU(h, cudaStreamSynchronize(h->st));  // caller buffers and local staging may go away now
  mark("copies complete");
  {
    const int rc = bind_problem(h, z);
    if (rc != DBA_OK) return rc;
  }
  h->keep.valid = h->world == 1;
  h->keep.xy = s_xy;
  h->keep.ip = s_ip;
  h->keep.ab = s_ab;
  h->have_problem = true;
  return dba_params_reset(h);
}

// Device-resident outer loop (reference sfm.cc:118-127: solve, filterPoint3d, solve, ...): drops the
// observations / points a dba_filter call flagged and rebuilds the index structures from the
// point-sorted arrays the engine kept from the last upload — the caller does not gather, validate,
// sort or upload the scene again, and the parameters continue from their current device values.
int dba_problem_update(dba_handle* h, const uint8_t* obs_remove, const uint8_t* pt_remove, int32_t freeze_camera,
                       int64_t* n_obs_out, int32_t* n_pts_out) {
  if (!h) return DBA_ERR_INVALID_ARGUMENT;
  if (!h->have_problem) return h->fail(DBA_ERR_NO_PROBLEM, "no problem set");
  if (h->world > 1) return h->fail(DBA_ERR_UNSUPPORTED, "dba_problem_update needs a single-GPU handle");
  CU(h, cudaSetDevice(h->device));
  if (!h->keep.valid && h->build_mode == 1) {
    // the device build keeps the point-sorted image on the device only: read it back once
    const size_t nk = static_cast<size_t>(h->n_obs);
    h->keep_xy_v.resize(nk);
    h->keep_ip_v.resize(nk);
    h->keep_ab_v.resize(nk);
    if (nk) {
      CU(h, cudaMemcpyAsync(h->keep_xy_v.data(), h->d_obs_xy.p, nk * sizeof(double2), cudaMemcpyDeviceToHost, h->st));
      CU(h, cudaMemcpyAsync(h->keep_ip_v.data(), h->d_obs_ip.p, nk * sizeof(int2), cudaMemcpyDeviceToHost, h->st));
      CU(h, cudaMemcpyAsync(h->keep_ab_v.data(), h->d_obs_ab.p, nk * sizeof(int2), cudaMemcpyDeviceToHost, h->st));
      CU(h, cudaStreamSynchronize(h->st));
    }
    h->keep.xy = h->keep_xy_v.data();
    h->keep.ip = h->keep_ip_v.data();
    h->keep.ab = h->keep_ab_v.data();
    h->keep.valid = true;
  }
  if (!h->keep.valid) return h->fail(DBA_ERR_NO_PROBLEM, "the retained problem image is gone: call dba_problem_set");
  const int64_t n = h->n_obs;
  const int n_pts = h->n_pts, n_ext = h->n_ext, n_intr = h->n_intr;
  const int c = h->cur;
  // current parameters: the scatter-back source, read once
  std::vector<double> pts(3 * static_cast<size_t>(n_pts)), rot(3 * static_cast<size_t>(n_ext)), trans(3 * static_cast<size_t>(n_ext)),
      focal(2 * static_cast<size_t>(n_intr)), dist(2 * static_cast<size_t>(n_intr));
  if (n_pts) CU(h, cudaMemcpyAsync(pts.data(), h->d_pts[c].p, pts.size() * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  if (n_ext) {
    CU(h, cudaMemcpyAsync(rot.data(), h->d_rot[c].p, rot.size() * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    CU(h, cudaMemcpyAsync(trans.data(), h->d_trans[c].p, trans.size() * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  }
  if (n_intr) {
    CU(h, cudaMemcpyAsync(focal.data(), h->d_focal[c].p, focal.size() * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    CU(h, cudaMemcpyAsync(dist.data(), h->d_dist[c].p, dist.size() * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  }
  // surviving points, re-indexed in order
  std::vector<int> new_pt(static_cast<size_t>(n_pts) + 1, 0);
  for (int i = 0; i < n_pts; ++i) new_pt[i + 1] = new_pt[i] + ((pt_remove && pt_remove[i]) ? 0 : 1);
  const int n_pts_new = new_pt[n_pts];
  // surviving observations in sorted order (two-pass chunked compaction on all host cores)
  const int2* ip = h->keep.ip;
  const int2* ab = h->keep.ab;
  const double2* xy = h->keep.xy;
  const int64_t* perm = h->perm.data();
  auto gone = [&](int64_t k) { return (obs_remove && obs_remove[perm[k]]) || (pt_remove && pt_remove[ip[k].y]); };
  const int n_chunk = std::max(1, omp_get_max_threads() * 4);
  std::vector<int64_t> first(static_cast<size_t>(n_chunk) + 1, 0);
#pragma omp parallel for schedule(static, 1)
  for (int ch = 0; ch < n_chunk; ++ch) {
    int64_t cnt = 0;
    for (int64_t k = n * ch / n_chunk; k < n * (ch + 1) / n_chunk; ++k) cnt += gone(k) ? 0 : 1;
    first[ch + 1] = cnt;
  }
  for (int ch = 0; ch < n_chunk; ++ch) first[ch + 1] += first[ch];
  const int64_t n_new = first[n_chunk];
  std::vector<double> nxy(2 * static_cast<size_t>(n_new));
  std::vector<int32_t> npt(static_cast<size_t>(n_new)), na(static_cast<size_t>(n_new)), nb(static_cast<size_t>(n_new)), nin(static_cast<size_t>(n_new));
  std::vector<int64_t> old_caller(static_cast<size_t>(n_new));
#pragma omp parallel for schedule(static, 1)
  for (int ch = 0; ch < n_chunk; ++ch) {
    int64_t w = first[ch];
    for (int64_t k = n * ch / n_chunk; k < n * (ch + 1) / n_chunk; ++k) {
      if (gone(k)) continue;
      nxy[2 * w] = xy[k].x;
      nxy[2 * w + 1] = xy[k].y;
      npt[w] = new_pt[ip[k].y];
      nin[w] = ip[k].x;
      na[w] = ab[k].x;
      nb[w] = ab[k].y;
      old_caller[w] = perm[k];
      ++w;
    }
  }
  // caller order of the survivors: rank of the old caller index among the kept ones
  const int64_t n_caller = h->n_obs_global;
  std::vector<int64_t> caller_rank(static_cast<size_t>(n_caller) + 1, 0);
  {
    std::vector<uint8_t> kept(static_cast<size_t>(n_caller), 0);
#pragma omp parallel for schedule(static)
    for (int64_t w = 0; w < n_new; ++w) kept[old_caller[w]] = 1;
    for (int64_t i = 0; i < n_caller; ++i) caller_rank[i + 1] = caller_rank[i] + kept[i];
  }
  CU(h, cudaStreamSynchronize(h->st));  // parameters have arrived
  std::vector<double> npts(3 * static_cast<size_t>(n_pts_new));
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n_pts; ++i)
    if (new_pt[i + 1] != new_pt[i])
      for (int k2 = 0; k2 < 3; ++k2) npts[3 * static_cast<size_t>(new_pt[i]) + k2] = pts[3 * static_cast<size_t>(i) + k2];
  // the small tables move out of the handle: dba_problem_set rewrites them
  std::vector<double> center = h->keep.center;
  std::vector<int32_t> nf = h->keep.nf, nd = h->keep.nd;
  std::vector<uint8_t> ext_const = h->keep.ext_const;
  dba_problem q;
  std::memset(&q, 0, sizeof q);
  q.n_obs = n_new;
  q.n_pts = n_pts_new;
  q.n_ext = n_ext;
  q.n_intr = n_intr;
  q.obs_xy = nxy.data();
  q.obs_pt = npt.data();
  q.obs_pose_a = na.data();
  q.obs_pose_b = nb.data();
  q.obs_intr = nin.data();
  q.pts = npts.data();
  q.ext_rot = rot.data();
  q.ext_trans = trans.data();
  q.intr_center = center.data();
  q.intr_focal = focal.data();
  q.intr_dist = dist.data();
  q.intr_nf = nf.data();
  q.intr_nd = nd.data();
  q.ext_const = ext_const.data();
  q.freeze_camera = freeze_camera;
  q.free_intrinsics = h->keep.free_intrinsics;
  const int rc = dba_problem_set(h, &q);
  if (rc != DBA_OK) return rc;
  // observation outputs (dba_eval, dba_filter) stay in the CALLER's order: position among the survivors
#pragma omp parallel for schedule(static)
  for (int64_t w = 0; w < n_new; ++w) h->perm[w] = caller_rank[old_caller[w]];
  h->n_obs_global = n_new;
  if (n_obs_out) *n_obs_out = n_new;
  if (n_pts_out) *n_pts_out = n_pts_new;
  return DBA_OK;
}

int dba_params_reset(dba_handle* h) {
  if (!h) return DBA_ERR_INVALID_ARGUMENT;
  if (!h->have_problem) return h->fail(DBA_ERR_NO_PROBLEM, "no problem set");
  CU(h, cudaSetDevice(h->device));
  h->cur = 0;
  auto cp = [&](DevBuf<double>* b, size_t count) -> cudaError_t {
    if (count == 0) return cudaSuccess;
    return cudaMemcpyAsync(b[0].p, b[2].p, count * sizeof(double), cudaMemcpyDeviceToDevice, h->st);
  };
  CU(h, cp(h->d_pts, 3 * static_cast<size_t>(h->n_pts)));
  CU(h, cp(h->d_rot, 3 * static_cast<size_t>(h->n_ext)));
  CU(h, cp(h->d_trans, 3 * static_cast<size_t>(h->n_ext)));
  CU(h, cp(h->d_focal, 2 * static_cast<size_t>(h->n_intr)));
  CU(h, cp(h->d_dist, 2 * static_cast<size_t>(h->n_intr)));
  CU(h, cudaStreamSynchronize(h->st));
  return DBA_OK;
}

// --------------------------------------------------------------------------- eval
int dba_eval(dba_handle* h, double* cost, double* residuals, double* jac_pt, double* jac_pose_a, double* jac_pose_b,
             double* jac_intr) {
  if (!h) return DBA_ERR_INVALID_ARGUMENT;
  if (!h->have_problem) return h->fail(DBA_ERR_NO_PROBLEM, "no problem set");
  CU(h, cudaSetDevice(h->device));
  const bool per_obs = residuals || jac_pt || jac_pose_a || jac_pose_b || jac_intr;
  if (per_obs && h->world > 1)
    return h->fail(DBA_ERR_UNSUPPORTED, "per-observation outputs need a single-GPU handle");
  h->D.loss_type = 0;  // dba_eval reports the raw functor (no loss), like Problem::Evaluate with apply_loss_function = false
  const DeviceProblem& D0 = h->D;
  const ParamSet& P = h->P[h->cur];
  const bool want_jac = jac_pt || jac_pose_a || jac_pose_b || jac_intr;
  {
    Scope s(h, "pose_rows");
    launch_pose_rows(P, h->d_ext_const.p, h->freeze, h->n_ext, h->n_intr, h->st);
  }
  const int64_t nl = h->n_obs;
  if (!want_jac) {
    DevBuf<double> mse;
    if (residuals) {
      // residual values come from the Jacobian kernel's r plane (same code path as the solver)
      DeviceProblem D = D0;
      DevBuf<double2> tmp;
      CU(h, tmp.alloc(D.ld * 4));
      D.J = tmp.p;
      {
        Scope s(h, "jacobian");
        launch_jacobian(D, P, h->W, 0, 0, 1, h->d_partA.p, h->st);
      }
      std::vector<double2> r(nl);
      CU(h, cudaMemcpyAsync(r.data(), tmp.p, nl * sizeof(double2), cudaMemcpyDeviceToHost, h->st));
      CU(h, cudaStreamSynchronize(h->st));
      for (int64_t k = 0; k < nl; ++k) {
        residuals[2 * h->perm[k]] = r[k].x;
        residuals[2 * h->perm[k] + 1] = r[k].y;
      }
    } else {
      Scope s(h, "cost", 24.0 * static_cast<double>(nl));
      launch_cost(D0, P, h->d_partA.p, nullptr, h->st);
    }
  } else {
    DeviceProblem D = D0;
    const int cbs = 9, twos = h->two;
    const int planes = 4 + cbs + (twos ? 6 : 0);
    DevBuf<double2> tmp;
    CU(h, tmp.alloc(D.ld * planes));
    D.J = tmp.p;
    {
      Scope s(h, "jacobian");
      launch_jacobian(D, P, h->W, cbs, twos, 1, h->d_partA.p, h->st);
    }
    std::vector<double2> host(static_cast<size_t>(D.ld) * planes);
    CU(h, cudaMemcpyAsync(host.data(), tmp.p, host.size() * sizeof(double2), cudaMemcpyDeviceToHost, h->st));
    CU(h, cudaStreamSynchronize(h->st));
    auto plane = [&](int pl, int64_t k) -> double2 { return host[static_cast<size_t>(pl) * D.ld + k]; };
    for (int64_t k = 0; k < nl; ++k) {
      const int64_t i = h->perm[k];
      if (residuals) {
        residuals[2 * i] = plane(kPlaneR, k).x;
        residuals[2 * i + 1] = plane(kPlaneR, k).y;
      }
      for (int c = 0; c < 3; ++c) {
        if (jac_pt) {
          jac_pt[6 * i + c] = plane(kPlaneJp + c, k).x;
          jac_pt[6 * i + 3 + c] = plane(kPlaneJp + c, k).y;
        }
        if (jac_intr) {
          jac_intr[6 * i + c] = plane(kPlaneJA + 6 + c, k).x;
          jac_intr[6 * i + 3 + c] = plane(kPlaneJA + 6 + c, k).y;
        }
      }
      for (int c = 0; c < 6; ++c) {
        if (jac_pose_a) {
          jac_pose_a[12 * i + c] = plane(kPlaneJA + c, k).x;
          jac_pose_a[12 * i + 6 + c] = plane(kPlaneJA + c, k).y;
        }
        if (jac_pose_b) {
          jac_pose_b[12 * i + c] = twos ? plane(kPlaneJA + 9 + c, k).x : 0.0;
          jac_pose_b[12 * i + 6 + c] = twos ? plane(kPlaneJA + 9 + c, k).y : 0.0;
        }
      }
    }
  }
  if (cost) {
    {
      Scope s(h, "reduce");
      // the cost partials come from k_cost (one per 256 observations) or from the Jacobian kernel
      launch_reduce_sum(h->d_partA.p, (!want_jac && !residuals) ? cost_grid(D0) : jacobian_partials(D0), 1, 0, h->W.scalars + S_COST, h->st);
    }
    int rc = reduce_scalars_and_fetch(h);
    if (rc != DBA_OK) return rc;
    *cost = 0.5 * h->h_scalars[S_COST];
  }
  CU(h, cudaGetLastError());
  return DBA_OK;
}

int dba_filter_mse(dba_handle* h, double* mse) {
  if (!h || !mse) return DBA_ERR_INVALID_ARGUMENT;
  if (!h->have_problem) return h->fail(DBA_ERR_NO_PROBLEM, "no problem set");
  if (h->world > 1) return h->fail(DBA_ERR_UNSUPPORTED, "per-observation outputs need a single-GPU handle");
  CU(h, cudaSetDevice(h->device));
  h->D.loss_type = 0;  // the filter re-evaluates the plain functor (DeepArcManager.cc:335-347)
  const ParamSet& P = h->P[h->cur];
  DevBuf<double> d;
  CU(h, d.alloc(std::max<int64_t>(h->n_obs, 1)));
  {
    Scope s(h, "pose_rows");
    launch_pose_rows(P, h->d_ext_const.p, h->freeze, h->n_ext, h->n_intr, h->st);
  }
  {
    Scope s(h, "cost", 24.0 * static_cast<double>(h->n_obs));
    launch_cost(h->D, P, h->d_partA.p, d.p, h->st);
  }
  std::vector<double> host(h->n_obs);
  CU(h, cudaMemcpyAsync(host.data(), d.p, h->n_obs * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  CU(h, cudaStreamSynchronize(h->st));
  for (int64_t k = 0; k < h->n_obs; ++k) mse[h->perm[k]] = host[k];
  return DBA_OK;
}

int dba_filter(dba_handle* h, double error_boundary, const double* centre, double rho, uint8_t* obs_remove, uint8_t* pt_remove,
               int64_t* n_obs_removed, int32_t* n_pts_removed) {
  if (!h || !obs_remove || !pt_remove) return DBA_ERR_INVALID_ARGUMENT;
  if (!h->have_problem) return h->fail(DBA_ERR_NO_PROBLEM, "no problem set");
  if (h->world > 1) return h->fail(DBA_ERR_UNSUPPORTED, "per-observation outputs need a single-GPU handle");
  CU(h, cudaSetDevice(h->device));
  h->D.loss_type = 0;
  const ParamSet& P = h->P[h->cur];
  const int64_t n = h->n_obs;
  const int n_pts = h->n_pts;
  DevBuf<double> d_mse;
  DevBuf<uint8_t> d_obs, d_pt;
  CU(h, d_mse.alloc(std::max<int64_t>(n, 1)));
  CU(h, d_obs.alloc(std::max<int64_t>(n, 1)));
  CU(h, d_pt.alloc(std::max(n_pts, 1)));
  {
    Scope s(h, "pose_rows");
    launch_pose_rows(P, h->d_ext_const.p, h->freeze, h->n_ext, h->n_intr, h->st);
  }
  {
    Scope s(h, "cost", 24.0 * static_cast<double>(n));
    launch_cost(h->D, P, h->d_partA.p, d_mse.p, h->st);
  }
  {
    Scope s(h, "filter_flags", 9.0 * static_cast<double>(n) + 25.0 * n_pts);
    launch_filter_flags(h->D, P, d_mse.p, error_boundary, centre, rho, d_obs.p, d_pt.p, h->st);
  }
  std::vector<uint8_t> sorted_flags(static_cast<size_t>(n));
  CU(h, cudaMemcpyAsync(sorted_flags.data(), d_obs.p, static_cast<size_t>(n), cudaMemcpyDeviceToHost, h->st));
  CU(h, cudaMemcpyAsync(pt_remove, d_pt.p, static_cast<size_t>(n_pts), cudaMemcpyDeviceToHost, h->st));
  CU(h, cudaStreamSynchronize(h->st));
  CU(h, cudaGetLastError());
  int64_t n_obs_gone = 0;
  const int64_t* perm = h->perm.data();
#pragma omp parallel for schedule(static) reduction(+ : n_obs_gone)
  for (int64_t k = 0; k < n; ++k) {
    obs_remove[perm[k]] = sorted_flags[k];
    n_obs_gone += sorted_flags[k];
  }
  int n_pts_gone = 0;
  for (int i = 0; i < n_pts; ++i) n_pts_gone += pt_remove[i];
  if (n_obs_removed) *n_obs_removed = n_obs_gone;
  if (n_pts_removed) *n_pts_removed = n_pts_gone;
  return DBA_OK;
}

// --------------------------------------------------------------------------- solve
void dba_solve_options_default(dba_solve_options* o) {
  if (!o) return;
  o->max_num_iterations = 50;
  o->max_solver_time_in_seconds = 1e9;
  o->initial_trust_region_radius = 1e4;
  o->max_trust_region_radius = 1e16;
  o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3;
  o->min_lm_diagonal = 1e-6;
  o->max_lm_diagonal = 1e32;
  o->function_tolerance = 1e-6;
  o->gradient_tolerance = 1e-10;
  o->parameter_tolerance = 1e-8;
  o->jacobi_scaling = 1;
  o->max_num_consecutive_invalid_steps = 5;
  o->linear_solver = DBA_LS_AUTO;
  o->pcg_max_iterations = 500;
  o->pcg_min_iterations = 0;
  o->pcg_rel_tolerance = 1e-12;
  o->dense_max_size = 768;
  o->progress_to_stdout = 0;
  o->loss_type = DBA_LOSS_NONE;
  o->loss_scale = 0.5;  // the value in the reference's commented-out CauchyLoss (sfm.cc:49)
}

int dba_solve(dba_handle* h, const dba_solve_options* opt, dba_summary* sum) {
  if (!h || !opt || !sum) return DBA_ERR_INVALID_ARGUMENT;
  if (!h->have_problem) return h->fail(DBA_ERR_NO_PROBLEM, "no problem set");
  CU(h, cudaSetDevice(h->device));
  const dba_solve_options o = *opt;
  dba_iteration* it_buf = sum->iterations;
  const int it_cap = it_buf ? sum->iterations_capacity : 0;
  std::memset(sum, 0, sizeof *sum);
  sum->iterations = it_buf;
  sum->iterations_capacity = it_cap;
  sum->termination = DBA_FAILURE;  // until finish() says otherwise: an error return never reads as converged
  sum->reduced_system_size = h->n_ext * h->cb;
  // linear solver: DENSE = the reference's DENSE_SCHUR (sfm.cc:67) as an explicit reduced system +
  // Cholesky; AUTO takes it whenever the reduced system is small enough, else the implicit PCG
  h->use_dense = false;
  if (h->cb) {
    if (o.linear_solver == DBA_LS_DENSE) {
      if (!h->dense_ok)
        return h->fail(DBA_ERR_UNSUPPORTED,
                       "DBA_LS_DENSE: the reduced system (%d camera blocks, %d unknowns) exceeds the dense path (%d blocks, %d unknowns); use DBA_LS_PCG or DBA_LS_AUTO",
                       h->n_ext, h->n_ext * h->cb, kDnMaxBlocks, kDnMaxSize);
      h->use_dense = true;
    } else if (o.linear_solver == DBA_LS_AUTO) {
      h->use_dense = h->dense_ok && h->n_ext * h->cb <= o.dense_max_size;
    } else if (o.linear_solver != DBA_LS_PCG) {
      return h->fail(DBA_ERR_INVALID_ARGUMENT, "unknown linear_solver %d", o.linear_solver);
    }
  }
  sum->linear_solver_used = (h->use_dense || !h->cb) ? DBA_LS_DENSE : DBA_LS_PCG;
  // robust loss (reference sfm.cc:49, commented out there): applied inside the Jacobian / cost kernels
  if (o.loss_type != DBA_LOSS_NONE && o.loss_type != DBA_LOSS_CAUCHY) return h->fail(DBA_ERR_INVALID_ARGUMENT, "unknown loss_type %d", o.loss_type);
  if (o.loss_type == DBA_LOSS_CAUCHY && !(o.loss_scale > 0.0)) return h->fail(DBA_ERR_INVALID_ARGUMENT, "loss_scale must be positive");
  h->D.loss_type = o.loss_type;
  h->D.loss_b = o.loss_scale * o.loss_scale;
  h->D.loss_c = o.loss_type ? 1.0 / h->D.loss_b : 0.0;
  // the matrix-free product recomputes the UNcorrected Jacobian from the camera rows: with a loss
  // the product reads the (corrected) planes instead
  struct MfGuard {
    dba_handle* h;
    int saved;
    ~MfGuard() { h->mf = saved; }
  } mf_guard{h, h->mf};
  if (o.loss_type != DBA_LOSS_NONE) h->mf = 0;
  h->planeless = h->mf && h->mf_front && h->cb > 0 && !h->two && !h->use_dense;
  h->dense_failures = 0;
  h->pcg_unconverged = 0;
  h->pcg_pending = false;
  const int64_t launches0 = h->launches;
  const double t_start = now_s();
  for (cudaEvent_t& e : h->ev_solve)
    if (!e) CU(h, cudaEventCreate(&e));
  const cudaEvent_t ev0 = h->ev_solve[0], ev1 = h->ev_solve[1], ev_loop = h->ev_solve[2];
  CU(h, cudaEventRecord(ev0, h->st));
  bool loop_started = false;

  int n_it = 0;
  auto push = [&](const dba_iteration& it) {
    if (n_it < it_cap) it_buf[n_it] = it;
    ++n_it;
  };
  double x_cost = 0.0;
  auto finish = [&](int termination, const char* msg) -> int {
    cudaEventRecord(ev1, h->st);
    cudaEventSynchronize(ev1);
    float ms = 0.f, ms_loop = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    if (loop_started) cudaEventElapsedTime(&ms_loop, ev_loop, ev1);
    sum->loop_device_time_in_seconds = ms_loop * 1e-3;
    sum->termination = termination;
    sum->num_iterations = std::min(n_it, it_cap);
    sum->final_cost = x_cost;
    sum->total_time_in_seconds = now_s() - t_start;
    sum->device_time_in_seconds = ms * 1e-3;
    sum->kernel_launches = h->launches - launches0;
    if (h->W.trace) {
      unsigned long long tr[16];
      if (cudaMemcpy(tr, h->W.trace, sizeof tr, cudaMemcpyDeviceToHost) == cudaSuccess && tr[0] > 0) {
        const double n = static_cast<double>(tr[0]);
        std::fprintf(stderr,
                     "[dba tail trace] rank %d launches %llu  us/launch: product %.1f | sync1 %.1f (CTA spread %.1f) | sums+push %.1f | "
                     "exchange %.1f | sync2 %.1f | phase2 %.1f | sync3 %.1f | phase3 %.1f\n",
                     h->rank, tr[0], tr[1] / n * 1e-3, tr[2] / n * 1e-3, tr[9] / n * 1e-3, tr[3] / n * 1e-3, tr[4] / n * 1e-3, tr[5] / n * 1e-3,
                     tr[6] / n * 1e-3, tr[7] / n * 1e-3, tr[8] / n * 1e-3);
        unsigned long long init[16] = {0};
        init[10] = ~0ull;
        cudaMemcpy(h->W.trace, init, sizeof init, cudaMemcpyHostToDevice);
      }
    }
    sum->linear_solver_failures = h->dense_failures;
    sum->pcg_unconverged_solves = h->pcg_unconverged;
    std::snprintf(sum->message, sizeof sum->message, "%s", msg);
    return DBA_OK;
  };

  if (h->n_obs_global == 0 || (h->n_pts_global == 0)) return finish(DBA_CONVERGENCE, "Function tolerance reached. No non-constant parameter blocks found.");

  h->jacobian_at_candidate = false;
  int rc = evaluate_jacobian(h, /*first=*/true, o.jacobi_scaling != 0);
  if (rc != DBA_OK) return rc;
  sum->jacobian_evaluations += o.jacobi_scaling ? 2 : 1;
  // speculate while steps keep being accepted (rejections cluster: after one, evaluate candidates
  // with the cheap residual-only kernel until a step succeeds again)
  bool expect_success = true;
  double radius = o.initial_trust_region_radius;
  double decrease_factor = 2.0;
  if ((rc = prepare_step(h, radius, o)) != DBA_OK) return rc;
  if ((rc = reduce_scalars_and_fetch(h)) != DBA_OK) return rc;
  const double* S = h->h_scalars;
  x_cost = 0.5 * S[S_COST];
  sum->initial_cost = x_cost;
  if (!std::isfinite(x_cost)) {
    finish(DBA_FAILURE, "Initial residual and Jacobian evaluation failed.");
    return h->fail(DBA_ERR_NUMERIC, "non-finite cost at the initial point");
  }
  dba_iteration it{};
  it.iteration = 0;
  it.cost = x_cost;
  it.gradient_max_norm = std::max(S[S_GMAX_PT], S[S_GMAX_CAM]);
  it.gradient_norm = std::sqrt(S[S_GSQ_PT] + S[S_GSQ_CAM]);
  bool linear_failure = (S[S_BAD_PT] + S[S_BAD_CAM]) > 0.0;
  int num_consecutive_invalid = 0;
  double iter_start = t_start;
  bool first_finalize = true;
  if (o.progress_to_stdout && h->rank == 0) print_progress_header();

  // FinalizeIterationAndCheckIfMinimizerCanContinue; returns 0 to continue, 1 when finished
  auto finalize = [&]() -> int {
    if (!first_finalize) {
      if (it.step_is_successful)
        sum->num_successful_steps++;
      else
        sum->num_unsuccessful_steps++;
    }
    first_finalize = false;
    it.trust_region_radius = radius;
    const double t = now_s();
    it.iteration_time_in_seconds = t - iter_start;
    push(it);
    if (o.progress_to_stdout && h->rank == 0) {
      std::printf("%4d % 8e   % 3.2e   % 3.2e  % 3.2e  % 3.2e % 3.2e     % 4d   % 3.2e   % 3.2e\n", it.iteration, it.cost,
                  it.cost_change, it.gradient_max_norm, it.step_norm, it.relative_decrease, it.trust_region_radius,
                  it.linear_solver_iterations, it.iteration_time_in_seconds, t - t_start);
      std::fflush(stdout);
    }
    if (t - t_start >= o.max_solver_time_in_seconds) {
      finish(DBA_NO_CONVERGENCE, "Maximum solver time reached.");
      return 1;
    }
    if (it.iteration >= o.max_num_iterations) {
      finish(DBA_NO_CONVERGENCE, "Maximum number of iterations reached.");
      return 1;
    }
    if (it.gradient_max_norm <= o.gradient_tolerance) {
      finish(DBA_CONVERGENCE, "Gradient tolerance reached.");
      return 1;
    }
    if (radius <= o.min_trust_region_radius) {
      finish(DBA_CONVERGENCE, "Minimum trust region radius reached.");
      return 1;
    }
    return 0;
  };

  char msg[192];
  while (finalize() == 0) {
    if (!loop_started) {
      CU(h, cudaEventRecord(ev_loop, h->st));
      loop_started = true;
    }
    iter_start = now_s();
    const dba_iteration prev = it;
    it = dba_iteration{};
    it.iteration = prev.iteration + 1;

    // ---- ComputeTrustRegionStep: Schur-complement solve on the device
    int pcg_iters = 0;
    bool step_valid = false;
    double model_cost_change = 0.0;
    if (!linear_failure) {
      if (h->cb) {
        if ((rc = h->use_dense ? dense_solve(h, &pcg_iters) : pcg_solve(h, o, &pcg_iters)) != DBA_OK) return rc;
      }
      const bool spec = h->speculate && expect_success;
      if ((rc = apply_step_and_evaluate(h, spec)) != DBA_OK) return rc;
      sum->residual_evaluations++;
      if (spec) sum->jacobian_evaluations++;
      if ((rc = reduce_scalars_and_fetch(h)) != DBA_OK) return rc;
      if ((rc = pcg_collect(h, &pcg_iters)) != DBA_OK) return rc;
      sum->pcg_iterations_total += pcg_iters;
      model_cost_change = -S[S_MODEL];
      step_valid = std::isfinite(model_cost_change) && model_cost_change > 0.0 &&
                   std::isfinite(S[S_STEP_PT] + S[S_STEP_CAM]);
    }
    it.linear_solver_iterations = h->cb ? pcg_iters : 1;
    it.model_cost_change = model_cost_change;
    it.step_is_valid = step_valid;
    if (!step_valid) {
      // HandleInvalidStep
      if (++num_consecutive_invalid >= o.max_num_consecutive_invalid_steps) {
        std::snprintf(msg, sizeof msg,
                      "Number of consecutive invalid steps more than "
                      "Solver::Options::max_num_consecutive_invalid_steps: %d",
                      o.max_num_consecutive_invalid_steps);
        it.cost = x_cost;
        it.trust_region_radius = radius;
        push(it);
        return finish(DBA_FAILURE, msg);
      }
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      it.cost = x_cost;
      it.gradient_max_norm = prev.gradient_max_norm;
      it.gradient_norm = prev.gradient_norm;
      expect_success = false;
      if ((rc = restore_current_jacobian(h, o.jacobi_scaling != 0)) != DBA_OK) return rc;
      if ((rc = prepare_step(h, radius, o)) != DBA_OK) return rc;
      if ((rc = reduce_scalars_and_fetch(h)) != DBA_OK) return rc;
      linear_failure = (S[S_BAD_PT] + S[S_BAD_CAM]) > 0.0;
      continue;
    }
    num_consecutive_invalid = 0;
    double candidate_cost = 0.5 * S[S_CAND];
    if (!std::isfinite(candidate_cost)) candidate_cost = std::numeric_limits<double>::max();
    const double x_norm = std::sqrt(S[S_X_PT] + S[S_X_CAM]);
    it.step_norm = std::sqrt(S[S_STEP_PT] + S[S_STEP_CAM]);
    // ---- ParameterToleranceReached
    const double step_size_tolerance = o.parameter_tolerance * (x_norm + o.parameter_tolerance);
    if (it.step_norm <= step_size_tolerance) {
      std::snprintf(msg, sizeof msg, "Parameter tolerance reached. Relative step_norm: %e <= %e.",
                    it.step_norm / (x_norm + o.parameter_tolerance), o.parameter_tolerance);
      it.cost = x_cost;
      it.trust_region_radius = radius;
      push(it);
      return finish(DBA_CONVERGENCE, msg);
    }
    // ---- FunctionToleranceReached
    it.cost_change = x_cost - candidate_cost;
    if (std::fabs(it.cost_change) <= o.function_tolerance * x_cost) {
      std::snprintf(msg, sizeof msg, "Function tolerance reached. |cost_change|/cost: %e <= %e",
                    std::fabs(it.cost_change) / x_cost, o.function_tolerance);
      it.cost = x_cost;
      it.trust_region_radius = radius;
      push(it);
      return finish(DBA_CONVERGENCE, msg);
    }
    // ---- IsStepSuccessful
    it.relative_decrease = candidate_cost >= std::numeric_limits<double>::max()
                               ? std::numeric_limits<double>::lowest()
                               : (x_cost - candidate_cost) / model_cost_change;
    if (it.relative_decrease > o.min_relative_decrease) {
      // HandleSuccessfulStep: the candidate becomes x; Jacobian at the new point
      h->cur = 1 - h->cur;
      if (h->jacobian_at_candidate) {
        if ((rc = adopt_candidate_jacobian(h)) != DBA_OK) return rc;
      } else {
        if ((rc = evaluate_jacobian(h, false, o.jacobi_scaling != 0)) != DBA_OK) return rc;
        sum->jacobian_evaluations++;
      }
      expect_success = true;
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * it.relative_decrease - 1.0, 3));
      radius = std::min(o.max_trust_region_radius, radius);
      decrease_factor = 2.0;
      if ((rc = prepare_step(h, radius, o)) != DBA_OK) return rc;
      if ((rc = reduce_scalars_and_fetch(h)) != DBA_OK) return rc;
      x_cost = 0.5 * S[S_COST];
      it.cost = x_cost;
      it.gradient_max_norm = std::max(S[S_GMAX_PT], S[S_GMAX_CAM]);
      it.gradient_norm = std::sqrt(S[S_GSQ_PT] + S[S_GSQ_CAM]);
      it.step_is_successful = 1;
    } else {
      it.step_is_successful = 0;
      it.cost = candidate_cost;
      it.gradient_max_norm = prev.gradient_max_norm;
      it.gradient_norm = prev.gradient_norm;
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      expect_success = false;
      if ((rc = restore_current_jacobian(h, o.jacobi_scaling != 0)) != DBA_OK) return rc;
      if ((rc = prepare_step(h, radius, o)) != DBA_OK) return rc;
      if ((rc = reduce_scalars_and_fetch(h)) != DBA_OK) return rc;
    }
    linear_failure = (S[S_BAD_PT] + S[S_BAD_CAM]) > 0.0;
  }
  return DBA_OK;
}

// ---------------------------------------------------------------------- params out
int dba_params_get(dba_handle* h, double* pts, double* ext_rot, double* ext_trans, double* intr_focal,
                   double* intr_dist) {
  if (!h) return DBA_ERR_INVALID_ARGUMENT;
  if (!h->have_problem) return h->fail(DBA_ERR_NO_PROBLEM, "no problem set");
  CU(h, cudaSetDevice(h->device));
  const int c = h->cur;
  if (pts) {
    // device -> pinned arena (one asynchronous copy at link rate) -> the caller's buffer (all host cores)
    const size_t total = 3 * static_cast<size_t>(h->world == 1 ? h->n_pts : h->n_pts_global);
    OmpThreadScope omp_scope(h->world);  // several ranks on one host: cores / world threads each, not cores each
    arena_of(h);
    PinnedArena& A = h->upload->readback;
    CU(h, cudaStreamSynchronize(h->st));
    CU(h, A.reserve(std::max(A.cap, total * sizeof(double) + 512)));
    double* stage = A.take<double>(std::max<size_t>(total, 1));
    if (h->world == 1) {
      CU(h, cudaMemcpyAsync(stage, h->d_pts[c].p, total * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    } else {
      // every rank returns the full array: zero-padded local slice, summed over ranks
      CU(h, ensure(h->d_full_pts, total));  // grow-only: cudaMalloc is slow once peer windows are mapped
      CU(h, cudaMemsetAsync(h->d_full_pts.p, 0, total * sizeof(double), h->st));
      CU(h, cudaMemcpyAsync(h->d_full_pts.p + 3 * static_cast<size_t>(h->pt_lo), h->d_pts[c].p,
                            3 * sizeof(double) * h->n_pts, cudaMemcpyDeviceToDevice, h->st));
      int rc = allreduce(h, h->d_full_pts.p, total, kNcclSum);
      if (rc != DBA_OK) return rc;
      CU(h, cudaMemcpyAsync(stage, h->d_full_pts.p, total * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    }
    CU(h, cudaStreamSynchronize(h->st));
#pragma omp parallel for schedule(static)
    for (int64_t ch = 0; ch < static_cast<int64_t>((total + 65535) / 65536); ++ch) {
      const size_t a = static_cast<size_t>(ch) * 65536, b = std::min(total, a + 65536);
      std::memcpy(pts + a, stage + a, (b - a) * sizeof(double));
    }
  }
  if (ext_rot) CU(h, cudaMemcpyAsync(ext_rot, h->d_rot[c].p, 3 * sizeof(double) * h->n_ext, cudaMemcpyDeviceToHost, h->st));
  if (ext_trans) CU(h, cudaMemcpyAsync(ext_trans, h->d_trans[c].p, 3 * sizeof(double) * h->n_ext, cudaMemcpyDeviceToHost, h->st));
  if (intr_focal) CU(h, cudaMemcpyAsync(intr_focal, h->d_focal[c].p, 2 * sizeof(double) * h->n_intr, cudaMemcpyDeviceToHost, h->st));
  if (intr_dist) CU(h, cudaMemcpyAsync(intr_dist, h->d_dist[c].p, 2 * sizeof(double) * h->n_intr, cudaMemcpyDeviceToHost, h->st));
  CU(h, cudaStreamSynchronize(h->st));
  return DBA_OK;
}

int dba_fit_hemisphere(dba_handle* h, const double* centres, int32_t n, double centre_io[3], double* rho_io,
                       const dba_solve_options* o, dba_summary* s) {
  if (!h || !centre_io || !rho_io || !o || !s || n < 0 || (n > 0 && !centres)) return DBA_ERR_INVALID_ARGUMENT;
  CU(h, cudaSetDevice(h->device));
  std::string err;
  int64_t launches = 0;
  int rc = hemisphere_fit_device(centres, n, centre_io, rho_io, o, s, h->st, &err, &launches);
  h->launches += launches;
  if (rc != DBA_OK) return h->fail(rc, "%s", err.c_str());
  return DBA_OK;
}

}  // extern "C"
