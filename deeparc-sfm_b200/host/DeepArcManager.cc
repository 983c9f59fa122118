// See DeepArcManager.hh.  Behavioural contract = reference src/DeepArcManager.cc; every
// function names the lines it mirrors.  Quirks that a drop-in must keep are marked QUIRK.
#include "DeepArcManager.hh"

#include <omp.h>

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <unordered_map>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "ba_client.hh"
#include "rotation_conv.hh"

namespace {

// Whitespace-separated token reader over the whole file image (the format is plain text:
// DeepArcManager.cc:36-48, :81, :103-118, :133-147, :159).  strtod/strtol on one buffer is
// ~10x faster than iostream extraction, which matters at 5-50 M observation lines.
class Tokens {
 public:
  explicit Tokens(const char* data) : p_(data) {}  // NUL-terminated image; the caller keeps it alive
  const char* pos() const { return p_; }
  void seek(const char* p) { p_ = p; }
  bool nextDouble(double* v) {
    skip();
    if (!*p_) return false;
    char* end = nullptr;
    *v = std::strtod(p_, &end);
    if (end == p_) return false;
    p_ = end;
    return true;
  }
  // istream >> int semantics: parse an integer prefix of the token
  bool nextInt(int* v) {
    skip();
    if (!*p_) return false;
    char* end = nullptr;
    const long x = std::strtol(p_, &end, 10);
    if (end == p_) return false;
    *v = static_cast<int>(x);
    p_ = end;
    return true;
  }

 private:
  void skip() {
    while (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r' || *p_ == '\f' || *p_ == '\v') ++p_;
  }
  const char* p_;
};

inline bool is_ws(char c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r' || c == '\f' || c == '\v'; }

// Start of every record of `stride` whitespace-separated tokens in [b, e): rec[r] = first byte of
// token r * stride for r = 0 .. n_rec (entry n_rec = the token after the last record, or e).
// Two parallel passes over the bytes (count token starts per chunk, then place).  Returns false
// when the text holds fewer than n_rec * stride tokens.
bool record_starts(const char* b, const char* e, int64_t n_rec, int stride, std::vector<const char*>* rec) {
  rec->assign(static_cast<size_t>(n_rec) + 1, nullptr);
  const int64_t len = e - b;
  const int n_chunk = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(omp_get_max_threads() * 8, len / 65536 + 1)));
  std::vector<int64_t> first(static_cast<size_t>(n_chunk) + 1, 0);
  auto starts_here = [b](const char* p) { return !is_ws(*p) && (p == b || is_ws(p[-1])); };
#pragma omp parallel for schedule(dynamic, 1)
  for (int c = 0; c < n_chunk; ++c) {
    const char* p = b + len * c / n_chunk;
    const char* q = b + len * (c + 1) / n_chunk;
    int64_t n = 0;
    for (; p < q; ++p) n += starts_here(p);
    first[c + 1] = n;
  }
  for (int c = 0; c < n_chunk; ++c) first[c + 1] += first[c];
  const int64_t need = n_rec * stride;
  if (first[n_chunk] < need) return false;
  (*rec)[n_rec] = e;
#pragma omp parallel for schedule(dynamic, 1)
  for (int c = 0; c < n_chunk; ++c) {
    int64_t g = first[c];
    if (g > need) continue;
    const char* p = b + len * c / n_chunk;
    const char* q = b + len * (c + 1) / n_chunk;
    for (; p < q && g <= need; ++p)
      if (starts_here(p)) {
        if (g % stride == 0) (*rec)[g / stride] = p;
        ++g;
      }
  }
  return true;
}

// strict token parsers for the parallel sections: the whole token must be a number (anything else
// sends the loader back to the serial tokenizer, which has istream's prefix semantics)
inline bool parse_int(const char*& p, const char* e, int* v) {
  while (p < e && is_ws(*p)) ++p;
  const char* s = p;
  if (s < e && *s == '+') ++s;
  auto r = std::from_chars(s, e, *v);
  if (r.ec != std::errc() || (r.ptr < e && !is_ws(*r.ptr))) return false;
  p = r.ptr;
  return true;
}
inline bool parse_double(const char*& p, const char* e, double* v) {
  while (p < e && is_ws(*p)) ++p;
  const char* s = p;
  if (s < e && *s == '+') ++s;
  auto r = std::from_chars(s, e, *v, std::chars_format::general);
  if (r.ec != std::errc() || (r.ptr < e && !is_ws(*r.ptr))) return false;
  p = r.ptr;
  return true;
}

// "%.6f".  Fast path: |v| < 1e6 and v * 1e6 further than 1e-3 from a rounding tie — the product
// carries an absolute error below 1.2e-4 there, so rounding it to an integer is the correctly
// rounded decimal printf produces; anything else goes through snprintf.
void fmt6(std::string* out, double v) {
  const double a = std::fabs(v);
  if (a < 1e6) {  // false for NaN
    const double q = a * 1e6;
    const double r = std::nearbyint(q);
    if (std::fabs(std::fabs(q - r) - 0.5) > 1e-3) {
      unsigned long long n = static_cast<unsigned long long>(r);
      char buf[24];
      int pos = 24;
      for (int i = 0; i < 6; ++i) {
        buf[--pos] = static_cast<char>('0' + n % 10);
        n /= 10;
      }
      buf[--pos] = '.';
      do {
        buf[--pos] = static_cast<char>('0' + n % 10);
        n /= 10;
      } while (n);
      if (std::signbit(v)) buf[--pos] = '-';
      out->append(buf + pos, static_cast<size_t>(24 - pos));
      return;
    }
  }
  char buf[400];  // DBL_MAX has 309 integer digits
  std::snprintf(buf, sizeof buf, "%.6f", v);
  out->append(buf);
}
void fmti(std::string* out, long long v) {
  char buf[24];
  auto r = std::to_chars(buf, buf + sizeof buf, v);
  out->append(buf, static_cast<size_t>(r.ptr - buf));
}
void fmtg(std::string* out, double v) {  // default ostream formatting of a double (precision 6)
  char buf[64];
  std::snprintf(buf, sizeof buf, "%g", v);
  out->append(buf);
}

}  // namespace

std::string deeparc_format_fixed6(double v) {
  std::string s;
  fmt6(&s, v);
  return s;
}

DeepArcManager::DeepArcManager() : arc_size_(0), ring_size_(0), share_extrinsic_(false) {}

DeepArcManager::~DeepArcManager() {
  if (deeparc::resident().manager == this) deeparc::resident_invalidate();  // the engine's copy describes objects that go away now
  for (ParameterBlock* b : params_) delete b;
  for (Point3d* p : point3d_) delete p;
  for (Intrinsic* i : intrinsics_) delete i;
  for (Extrinsic* e : extrinsics_) delete e;
  for (Camera* c : camera_) delete c;
  for (auto& row : hemisphere_)
    for (auto& cell : row.second) delete cell.second;
}

bool DeepArcManager::isShareExtrinsic() { return share_extrinsic_; }
std::vector<ParameterBlock*>* DeepArcManager::parameters() { return &params_; }
std::vector<Point3d*>* DeepArcManager::point3ds() { return &point3d_; }

// ring position -> slot in the extrinsic table; ring 0 aliases slot 0 (DeepArcManager.cc:166-171)
int DeepArcManager::ringSlot(int ring_position, int arc_size) {
  return ring_position == 0 ? 0 : ring_position + arc_size - 1;
}

// DeepArcManager.cc:26-74.  The two O(n) sections (observations, points) are parsed, allocated
// and linked by all host cores (at 5-50 M observation lines the reference's
// iostream loader takes longer than the GPU solve by two orders of magnitude); the small sections
// and any file the strict parallel parser does not accept go through the serial tokenizer.
namespace {
// The file image: mmap'ed (no copy; the kernel reads ahead while the parser threads run), or read
// into memory when the size is an exact number of pages (the text tokenizer wants a NUL after the
// last byte, which the zero-filled tail of the last mapped page provides otherwise).
struct FileImage {
  const char* data = nullptr;
  size_t size = 0;
  void* map = nullptr;
  size_t map_len = 0;
  std::string owned;
  bool open(const std::string& filename) {
    const int fd = ::open(filename.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (::fstat(fd, &st) != 0) {
      ::close(fd);
      return false;
    }
    size = static_cast<size_t>(st.st_size);
    const size_t page = static_cast<size_t>(::sysconf(_SC_PAGESIZE));
    if (size > 0 && size % page != 0) {
      void* m = ::mmap(nullptr, size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
      if (m != MAP_FAILED) {
        ::madvise(m, size, MADV_SEQUENTIAL | MADV_WILLNEED);
        map = m;
        map_len = size;
        data = static_cast<const char*>(m);
        ::close(fd);
        return true;
      }
    }
    owned.resize(size);
    size_t got = 0;
    while (got < size) {
      const ssize_t n = ::read(fd, &owned[got], size - got);
      if (n <= 0) break;
      got += static_cast<size_t>(n);
    }
    owned.resize(got);
    size = got;
    data = owned.c_str();
    ::close(fd);
    return true;
  }
  ~FileImage() {
    if (map) ::munmap(map, map_len);
  }
};

// ---- binary side format (".deeparcb").  The text format keeps six decimals
// (DeepArcManager.cc:428), so a text round trip is lossy; this one stores the scene bit for bit and
// loads at memcpy speed.  Little-endian, all sections 8-byte aligned:
//   char magic[8] = "DEEPARCB"; u32 version = 1; u32 reserved;
//   i64 n_obs; i32 n_intr, n_arc, n_ring (0: non-shared), n_pts, n_ext, pad;
//   i32 col0[n_obs] (pos_arc | intrinsic id)  i32 col1[n_obs] (pos_ring | extrinsic id)  i32 point[n_obs]  (+pad)
//   f64 xy[n_obs][2]
//   per intrinsic: f64 cx, cy, f0, f1, k0, k1; i32 nf, nd      per extrinsic: f64 t[3], aa[3]
//   f64 xyz[n_pts][3]; i32 rgb[n_pts][3] (+pad)
constexpr char kBinMagic[8] = {'D', 'E', 'E', 'P', 'A', 'R', 'C', 'B'};
struct BinHeader {
  char magic[8];
  uint32_t version, reserved;
  int64_t n_obs;
  int32_t n_intr, n_arc, n_ring, n_pts, n_ext, pad;
};
struct BinIntrinsic {
  double cx, cy, f0, f1, k0, k1;
  int32_t nf, nd;
};
struct BinExtrinsic {
  double t[3], aa[3];
};
inline size_t pad8(size_t n) { return (n + 7) & ~size_t{7}; }
}  // namespace

// DeepArcManager.cc:26-74.  The two O(n) sections (observations, points) are parsed, allocated
// and linked by all host cores (at 5-50 M observation lines the reference's
// iostream loader takes longer than the GPU solve by two orders of magnitude); the small sections
// and any file the strict parallel parser does not accept go through the serial tokenizer.
// A file that starts with the binary magic is loaded by readBinary.
bool DeepArcManager::read(std::string filename) {
  FileImage img;
  if (!img.open(filename)) {
    std::cout << "Cannot read " << filename << std::endl;
    throw "Cannot read input file";
  }
  if (img.size >= sizeof(BinHeader) && std::memcmp(img.data, kBinMagic, 8) == 0) {
    if (!readBinary(img.data, img.size)) throw "Corrupt binary deeparc file";
    return true;
  }
  const char* serial = std::getenv("DEEPARC_SERIAL_IO");
  const bool parallel = !(serial && serial[0] == '1');
  if (parallel && readText(img.data, img.size, true)) return true;
  if (parallel) clearScene();  // the strict parser met a token it does not take: start over
  readText(img.data, img.size, false);
  return true;
}

bool DeepArcManager::readBinary(const char* data, size_t size) {
  BinHeader h;
  std::memcpy(&h, data, sizeof h);
  if (h.version != 1 || h.n_obs < 0 || h.n_intr < 0 || h.n_pts < 0 || h.n_ext < 0) return false;
  const size_t n = static_cast<size_t>(h.n_obs);
  size_t off = sizeof(BinHeader);
  const size_t o_col0 = off, o_col1 = o_col0 + 4 * n, o_pid = o_col1 + 4 * n;
  off = pad8(o_pid + 4 * n);
  const size_t o_xy = off;
  off += 16 * n;
  const size_t o_intr = off;
  off += sizeof(BinIntrinsic) * static_cast<size_t>(h.n_intr);
  const size_t o_ext = off;
  off += sizeof(BinExtrinsic) * static_cast<size_t>(h.n_ext);
  const size_t o_xyz = off;
  off += 24 * static_cast<size_t>(h.n_pts);
  const size_t o_rgb = off;
  off += 12 * static_cast<size_t>(h.n_pts);
  if (off > size) return false;
  share_extrinsic_ = h.n_ring != 0;
  arc_size_ = h.n_arc;
  ring_size_ = h.n_ring;
  const int32_t* col0 = reinterpret_cast<const int32_t*>(data + o_col0);
  const int32_t* col1 = reinterpret_cast<const int32_t*>(data + o_col1);
  const int32_t* pid = reinterpret_cast<const int32_t*>(data + o_pid);
  const double* xy = reinterpret_cast<const double*>(data + o_xy);
  params_.assign(n, nullptr);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < h.n_obs; ++i) params_[i] = new ParameterBlock(col0[i], col1[i], pid[i], new Point2d(xy[2 * i], xy[2 * i + 1]));
  for (int i = 0; i < h.n_intr; ++i) {
    BinIntrinsic b;
    std::memcpy(&b, data + o_intr + sizeof b * static_cast<size_t>(i), sizeof b);
    Intrinsic* in = new Intrinsic();
    in->id(i);
    in->center(static_cast<int>(b.cx), static_cast<int>(b.cy));  // same entry point as the text loader (already integral)
    double f[2] = {b.f0, b.f1}, k[2] = {b.k0, b.k1};
    in->focal(b.nf, f);
    in->distrotion(b.nd, k);
    intrinsics_.push_back(in);
  }
  for (int i = 0; i < h.n_ext; ++i) {
    BinExtrinsic b;
    std::memcpy(&b, data + o_ext + sizeof b * static_cast<size_t>(i), sizeof b);
    Extrinsic* ex = new Extrinsic();
    ex->id(i);
    ex->translation(b.t[0], b.t[1], b.t[2]);
    ex->rotation(b.aa);
    extrinsics_.push_back(ex);
  }
  const double* xyz = reinterpret_cast<const double*>(data + o_xyz);
  const int32_t* rgb = reinterpret_cast<const int32_t*>(data + o_rgb);
  point3d_.assign(static_cast<size_t>(h.n_pts), nullptr);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < h.n_pts; ++i)
    point3d_[i] = new Point3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
  if (share_extrinsic_)
    buildHemisphere();
  else
    buildCameras();
  linkBlocks(share_extrinsic_ ? arc_size_ : 0);
  return true;
}

// Same scene as write(), bit for bit (the observation columns are the ones write() emits:
// intrinsic()->id() and ring()->id() / extrinsic()->id(), points re-indexed; DeepArcManager.cc:430-449).
void DeepArcManager::writeBinary(std::string filename) {
  const int64_t n_pts = static_cast<int64_t>(point3d_.size()), n_obs = static_cast<int64_t>(params_.size());
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n_pts; ++i) point3d_[i]->id(static_cast<int>(i));
  BinHeader h;
  std::memset(&h, 0, sizeof h);
  std::memcpy(h.magic, kBinMagic, 8);
  h.version = 1;
  h.n_obs = n_obs;
  h.n_intr = static_cast<int32_t>(intrinsics_.size());
  h.n_arc = share_extrinsic_ ? arc_size_ : static_cast<int32_t>(camera_.size());
  h.n_ring = share_extrinsic_ ? ring_size_ : 0;
  h.n_pts = static_cast<int32_t>(n_pts);
  h.n_ext = static_cast<int32_t>(extrinsics_.size());
  const size_t n = static_cast<size_t>(n_obs);
  std::vector<char> buf(sizeof h + pad8(12 * n) + 16 * n + sizeof(BinIntrinsic) * intrinsics_.size() +
                        sizeof(BinExtrinsic) * extrinsics_.size() + 24 * static_cast<size_t>(n_pts) + pad8(12 * static_cast<size_t>(n_pts)), 0);
  char* p = buf.data();
  std::memcpy(p, &h, sizeof h);
  int32_t* col0 = reinterpret_cast<int32_t*>(p + sizeof h);
  int32_t* col1 = col0 + n;
  int32_t* pid = col1 + n;
  double* xy = reinterpret_cast<double*>(p + sizeof h + pad8(12 * n));
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n_obs; ++i) {
    ParameterBlock* b = params_[i];
    col0[i] = b->intrinsic()->id();
    col1[i] = share_extrinsic_ ? b->ring()->id() : b->extrinsic()->id();
    pid[i] = b->point3d()->id();
    xy[2 * i] = b->point2d()->x();
    xy[2 * i + 1] = b->point2d()->y();
  }
  char* q = reinterpret_cast<char*>(xy + 2 * n);
  for (Intrinsic* in : intrinsics_) {
    BinIntrinsic b;
    std::memset(&b, 0, sizeof b);
    b.cx = in->center()[0];
    b.cy = in->center()[1];
    b.nf = in->focal_size();
    b.nd = in->distrotion_size();
    b.f0 = in->focal()[0];
    b.f1 = in->focal()[1];
    b.k0 = in->distrotion()[0];
    b.k1 = in->distrotion()[1];
    std::memcpy(q, &b, sizeof b);
    q += sizeof b;
  }
  for (Extrinsic* ex : extrinsics_) {
    BinExtrinsic b;
    for (int j = 0; j < 3; ++j) {
      b.t[j] = ex->translation()[j];
      b.aa[j] = ex->rotation()[j];
    }
    std::memcpy(q, &b, sizeof b);
    q += sizeof b;
  }
  double* xyz = reinterpret_cast<double*>(q);
  int32_t* rgb = reinterpret_cast<int32_t*>(q + 24 * static_cast<size_t>(n_pts));
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n_pts; ++i) {
    Point3d* pt = point3d_[i];
    for (int j = 0; j < 3; ++j) xyz[3 * i + j] = pt->position()[j];
    rgb[3 * i] = pt->r();
    rgb[3 * i + 1] = pt->g();
    rgb[3 * i + 2] = pt->b();
  }
  std::ofstream of(filename, std::ios::binary);
  of.write(buf.data(), static_cast<std::streamsize>(buf.size()));
}

void DeepArcManager::clearScene() {
  if (deeparc::resident().manager == this) deeparc::resident_invalidate();
  for (ParameterBlock* b : params_) {
    if (b) b->point3d_unlinked(nullptr);
    delete b;
  }
  for (Point3d* p : point3d_) delete p;
  for (Intrinsic* i : intrinsics_) delete i;
  for (Extrinsic* e : extrinsics_) delete e;
  for (Camera* c : camera_) delete c;
  for (auto& row : hemisphere_)
    for (auto& cell : row.second) delete cell.second;
  params_.clear();
  point3d_.clear();
  intrinsics_.clear();
  extrinsics_.clear();
  camera_.clear();
  hemisphere_.clear();
}

bool DeepArcManager::readText(const char* data, size_t size, bool parallel) {
  const bool timing = std::getenv("DBA_TIMING") != nullptr;
  double t_mark = omp_get_wtime();
  auto mark = [&](const char* what) {
    if (!timing) return;
    const double t = omp_get_wtime();
    std::fprintf(stderr, "[DeepArcManager::read] %-24s %8.2f ms\n", what, 1e3 * (t - t_mark));
    t_mark = t;
  };
  Tokens tok(data);
  const char* const end = data + size;

  double version = 0.0;
  int n_block = 0, n_intrinsic = 0, n_arc = 0, n_ring = 0, n_point = 0;
  tok.nextDouble(&version);
  tok.nextInt(&n_block);
  tok.nextInt(&n_intrinsic);
  tok.nextInt(&n_arc);
  tok.nextInt(&n_ring);
  tok.nextInt(&n_point);
  share_extrinsic_ = n_ring != 0;
  arc_size_ = n_arc;
  ring_size_ = n_ring;
  const int n_extrinsic = n_ring != 0 ? n_arc + n_ring - 1 : n_arc;

  // observations: pos_arc pos_ring point3d_id x y   (:76-91)
  const size_t first_block = params_.size();
  if (parallel && n_block > 0) {
    std::vector<const char*> rec;
    if (!record_starts(tok.pos(), end, n_block, 5, &rec)) return false;
    mark("observation record index");
    params_.resize(first_block + static_cast<size_t>(n_block), nullptr);
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int i = 0; i < n_block; ++i) {
      const char* p = rec[i];
      const char* e = rec[i + 1];
      int a = 0, r = 0, pid = 0;
      double x = 0.0, y = 0.0;
      if (parse_int(p, e, &a) && parse_int(p, e, &r) && parse_int(p, e, &pid) && parse_double(p, e, &x) && parse_double(p, e, &y))
        params_[first_block + i] = new ParameterBlock(a, r, pid, new Point2d(x, y));
      else
        bad |= 1;
    }
    if (bad) return false;
    tok.seek(rec[n_block]);
  } else {
    params_.reserve(params_.size() + static_cast<size_t>(std::max(n_block, 0)));
    for (int i = 0; i < n_block; ++i) {
      int a = 0, r = 0, pid = 0;
      double x = 0.0, y = 0.0;
      tok.nextInt(&a);
      tok.nextInt(&r);
      tok.nextInt(&pid);
      tok.nextDouble(&x);
      tok.nextDouble(&y);
      params_.push_back(new ParameterBlock(a, r, pid, new Point2d(x, y)));
    }
  }
  mark("observations");
  // intrinsics: cx cy nf f.. nd k..   (:93-122)
  std::vector<Intrinsic*> intrinsics;
  for (int i = 0; i < n_intrinsic; ++i) {
    Intrinsic* in = new Intrinsic();
    in->id(i);
    double cx = 0.0, cy = 0.0, f[2] = {0.0, 0.0}, k[2] = {0.0, 0.0}, hold = 0.0;
    int nf = 0, nd = 0;
    tok.nextDouble(&cx);
    tok.nextDouble(&cy);
    in->center(static_cast<int>(cx), static_cast<int>(cy));  // QUIRK: truncated principal point
    tok.nextInt(&nf);
    for (int j = 0; j < nf; ++j) {
      tok.nextDouble(&hold);
      if (j < 2) f[j] = hold;  // the reference overruns a 2-element buffer here; we drop extras
    }
    in->focal(nf, f);
    tok.nextInt(&nd);
    for (int j = 0; j < nd; ++j) {
      tok.nextDouble(&hold);
      if (j < 2) k[j] = hold;
    }
    in->distrotion(nd, k);
    intrinsics.push_back(in);
  }
  // extrinsics: tx ty tz nrot rot..   (:124-151); nrot 3 angle-axis, 4 quaternion wxyz,
  // 9 COLUMN-MAJOR matrix
  std::vector<Extrinsic*> extrinsics;
  for (int i = 0; i < n_extrinsic; ++i) {
    Extrinsic* ex = new Extrinsic();
    ex->id(i);
    double t[3] = {0, 0, 0}, rot[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, aa[3] = {0, 0, 0}, hold = 0.0;
    int nrot = 0;
    tok.nextDouble(&t[0]);
    tok.nextDouble(&t[1]);
    tok.nextDouble(&t[2]);
    ex->translation(t[0], t[1], t[2]);
    tok.nextInt(&nrot);
    for (int j = 0; j < nrot; ++j) {
      tok.nextDouble(&hold);
      if (j < 9) rot[j] = hold;
    }
    if (nrot == 9)
      deeparc::rotation_matrix_to_angle_axis(rot, aa);
    else if (nrot == 4)
      deeparc::quaternion_to_angle_axis(rot, aa);
    ex->rotation(nrot == 3 ? rot : aa);  // any other count: zero rotation (reference: uninitialised)
    extrinsics.push_back(ex);
  }
  intrinsics_ = intrinsics;
  extrinsics_ = extrinsics;
  // points: x y z r g b (colour parsed as double, stored as int; :153-164)
  std::vector<Point3d*> points;
  if (parallel && n_point > 0) {
    std::vector<const char*> rec;
    if (!record_starts(tok.pos(), end, n_point, 6, &rec)) return false;
    mark("point record index");
    points.assign(static_cast<size_t>(n_point), nullptr);
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int i = 0; i < n_point; ++i) {
      const char* p = rec[i];
      const char* e = rec[i + 1];
      double v[6] = {0, 0, 0, 0, 0, 0};
      bool ok = true;
      for (double& x : v) ok = ok && parse_double(p, e, &x);
      if (ok)
        points[i] = new Point3d(v[0], v[1], v[2], static_cast<int>(v[3]), static_cast<int>(v[4]), static_cast<int>(v[5]));
      else
        bad |= 1;
    }
    point3d_ = points;  // owned from here on (clearScene frees them if the strict parser gives up)
    if (bad) return false;
  } else {
    points.reserve(static_cast<size_t>(std::max(n_point, 0)));
    for (int i = 0; i < n_point; ++i) {
      double v[6] = {0, 0, 0, 0, 0, 0};
      for (double& x : v) tok.nextDouble(&x);
      points.push_back(new Point3d(v[0], v[1], v[2], static_cast<int>(v[3]), static_cast<int>(v[4]), static_cast<int>(v[5])));
    }
    point3d_ = points;
  }

  mark("points");
  if (share_extrinsic_)
    buildHemisphere();
  else
    buildCameras();
  mark("cameras");
  linkBlocks(share_extrinsic_ ? arc_size_ : 0);
  mark("link");
  return true;
}

// DeepArcManager.cc:173-196 — resolve ids to pointers; out-of-range ids throw std::out_of_range
void DeepArcManager::linkBlocks(int arc_size) {
  const int64_t n = static_cast<int64_t>(params_.size());
  const int n_in = static_cast<int>(intrinsics_.size()), n_ex = static_cast<int>(extrinsics_.size());
  const int n_pt = static_cast<int>(point3d_.size());
  // pass 1 (parallel): ids -> pointers; the first id out of range is reported like vector::at()
  int64_t first_bad = n;
#pragma omp parallel for schedule(static) reduction(min : first_bad)
  for (int64_t i = 0; i < n; ++i) {
    ParameterBlock* p = params_[i];
    const int ring_slot = arc_size != 0 ? ringSlot(p->pos_ring(), arc_size) : 0;
    const bool ok = p->intrinsic_id() >= 0 && p->intrinsic_id() < n_in && p->point3d_id() >= 0 && p->point3d_id() < n_pt &&
                    (arc_size != 0 ? (p->pos_arc() >= 0 && p->pos_arc() < n_ex && ring_slot >= 0 && ring_slot < n_ex)
                                   : (p->extrinsic_id() >= 0 && p->extrinsic_id() < n_ex));
    if (!ok) {
      first_bad = std::min(first_bad, i);
      continue;
    }
    p->intrinsic(intrinsics_[p->intrinsic_id()]);
    p->point3d_unlinked(point3d_[p->point3d_id()]);
    if (arc_size != 0) {
      p->arc(extrinsics_[p->pos_arc()]);
      p->ring(extrinsics_[ring_slot]);
      p->share_extrinsic(true);
    } else {
      p->extrinsic(extrinsics_[p->extrinsic_id()]);
      p->share_extrinsic(false);
    }
  }
  if (first_bad < n) {
    // blocks that never got a point must not unlink from one in their destructor
    for (int64_t i = 0; i < n; ++i) params_[i]->point3d_unlinked(nullptr);
    throw std::out_of_range("vector::_M_range_check: DeepArcManager::linkBlocks: id out of range in observation " +
                            std::to_string(first_bad));
  }
  // pass 2: back-links point -> observations, grouped by point without a lock
  std::vector<int> count(static_cast<size_t>(n_pt), 0), cursor(static_cast<size_t>(n_pt), 0);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
#pragma omp atomic
    count[params_[i]->point3d_id()]++;
  }
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n_pt; ++i) point3d_[i]->link_slots(static_cast<size_t>(count[i]));
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const int pid = params_[i]->point3d_id();
    int slot;
#pragma omp atomic capture
    slot = cursor[pid]++;
    point3d_[pid]->link_slot(static_cast<size_t>(slot), params_[i]);
  }
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n_pt; ++i) point3d_[i]->link_finish();
}

// DeepArcManager.cc:198-218.  QUIRK: extrinsic ids are overwritten with their arc / ring
// POSITION (slot 0 ends up with id 0 either way); write() emits those ids.
void DeepArcManager::buildHemisphere() {
  for (int a = 0; a < arc_size_; ++a) {
    extrinsics_.at(a)->id(a);
    for (int r = 0; r < ring_size_; ++r) {
      const int slot = ringSlot(r, arc_size_);
      extrinsics_.at(slot)->id(r);
      hemisphere_[a][r] = new Camera(intrinsics_.at(a), extrinsics_.at(a), extrinsics_.at(slot));
    }
  }
}

// DeepArcManager.cc:220-240 — one camera per distinct extrinsic id, first intrinsic seen wins
void DeepArcManager::buildCameras() {
  // extrinsic id -> first intrinsic id seen with it, visited in extrinsic-id order (the reference's
  // std::map<int,int>::insert per observation); ids inside the table use a flat array
  const int n_ex = static_cast<int>(extrinsics_.size());
  std::vector<int> first_intr(static_cast<size_t>(std::max(n_ex, 0)), -1);
  std::vector<char> seen(first_intr.size(), 0);
  std::map<int, int> outside;  // ids that at() will reject below, exactly like the reference
  for (ParameterBlock* p : params_) {
    const int e = p->extrinsic_id();
    if (e >= 0 && e < n_ex) {
      if (!seen[e]) {
        seen[e] = 1;
        first_intr[e] = p->intrinsic_id();
      }
    } else {
      outside.insert(std::make_pair(e, p->intrinsic_id()));
    }
  }
  std::map<int, int> intrinsic_of(outside);
  for (int e = 0; e < n_ex; ++e)
    if (seen[e]) intrinsic_of.insert(std::make_pair(e, first_intr[e]));
  for (const auto& kv : intrinsic_of) camera_.push_back(new Camera(intrinsics_.at(kv.second), extrinsics_.at(kv.first)));
}

// camera centre -R^T t (DeepArcManager.cc:242-251)
std::vector<double> DeepArcManager::centreOf(Extrinsic* pose) {
  double m[9], c[3];
  deeparc::angle_axis_to_rotation_matrix(pose->rotation(), m);
  deeparc::camera_centre(m, pose->translation(), c);
  return std::vector<double>(c, c + 3);
}

// composed pose x_cam = R_arc (R_ring x + t_ring) + t_arc:
// centre = -R_ring^T t_ring - R_ring^T R_arc^T t_arc (DeepArcManager.cc:253-264)
std::vector<double> DeepArcManager::centreOf(Extrinsic* arc, Extrinsic* ring) {
  // evaluated in the association of the reference's expression:
  //   (-R1^T) t1 - ((R1^T R2^T) t2),  R1 = ring, R2 = arc
  double m1[9], m2[9], c1[3];
  deeparc::angle_axis_to_rotation_matrix(ring->rotation(), m1);
  deeparc::angle_axis_to_rotation_matrix(arc->rotation(), m2);
  deeparc::camera_centre(m1, ring->translation(), c1);
  const double* t2 = arc->translation();
  std::vector<double> c(3);
  for (int i = 0; i < 3; ++i) {
    double acc = 0.0;
    for (int j = 0; j < 3; ++j) {
      // (R1^T R2^T)(i, j) = sum_k R1(k, i) R2(j, k); column-major storage m[col * 3 + row]
      double mij = 0.0;
      for (int k = 0; k < 3; ++k) mij += m1[i * 3 + k] * m2[k * 3 + j];
      acc += mij * t2[j];
    }
    c[i] = c1[i] - acc;
  }
  return c;
}

// which pose(s) define camera (arc, ring): same selection as ParameterBlock::get()
std::vector<double> DeepArcManager::centreOfCamera(int arc, int ring, bool* single_pose) {
  Camera* cam = hemisphere_[arc][ring];
  if (ring == 0) {
    *single_pose = true;
    return centreOf(cam->arc());
  }
  if (arc == 0) {
    *single_pose = true;
    return centreOf(cam->ring());
  }
  *single_pose = false;
  return centreOf(cam->arc(), cam->ring());
}

// DeepArcManager.cc:501-518 (empty for non-shared files: ring_size_ == 0)
std::vector<std::vector<double> > DeepArcManager::getCameraCenter() {
  std::vector<std::vector<double> > centres;
  for (int a = 0; a < arc_size_; ++a)
    for (int r = 0; r < ring_size_; ++r) {
      bool single = false;
      centres.push_back(centreOfCamera(a, r, &single));
    }
  return centres;
}

// ASCII PLY: camera centres (green = single pose, magenta = composed), then the points
// (DeepArcManager.cc:266-328).  Numbers use default ostream formatting (%g, 6 digits).
void DeepArcManager::writePly(std::string filename) {
  const int n_cam = share_extrinsic_ ? arc_size_ * ring_size_ : static_cast<int>(camera_.size());
  std::string out;
  out.reserve(64 * (point3d_.size() + static_cast<size_t>(n_cam)) + 256);
  out += "ply\nformat ascii 1.0\nelement vertex " + std::to_string(point3d_.size() + static_cast<size_t>(n_cam)) +
         "\nproperty float x\nproperty float y\nproperty float z\nproperty uchar red\nproperty uchar green\n"
         "property uchar blue\nend_header\n";
  auto emit_cam = [&out](const std::vector<double>& c, bool green) {
    for (int i = 0; i < 3; ++i) {
      fmtg(&out, c[i]);
      out += ' ';
    }
    out += green ? "0 255 0\n" : "255 0 255\n";
  };
  if (share_extrinsic_) {
    for (int a = 0; a < arc_size_; ++a)
      for (int r = 0; r < ring_size_; ++r) {
        bool single = false;
        const std::vector<double> c = centreOfCamera(a, r, &single);
        emit_cam(c, single);
      }
  } else {
    for (Camera* cam : camera_) emit_cam(centreOf(cam->extrinsic()), true);
  }
  for (Point3d* p : point3d_) {
    const double* x = p->position();
    for (int j = 0; j < 3; ++j) {
      fmtg(&out, x[j]);
      out += ' ';
    }
    out += std::to_string(p->r()) + ' ' + std::to_string(p->g()) + ' ' + std::to_string(p->b()) + '\n';
  }
  std::ofstream of(filename, std::ios::binary);
  of.write(out.data(), static_cast<std::streamsize>(out.size()));
}

// DeepArcManager.cc:331-424.  The per-observation residuals (:332-352) are evaluated by the
// GPU engine at the parameters currently in the scene graph; the pointer-graph surgery that
// follows is the reference's, step for step.
void DeepArcManager::filterPoint3d(double error_boundary, double* hemisphere_center, double hemisphere_radius) {
  // The three removal rules (:347-350 mse < boundary, :368-378 points left empty, :380-408 points
  // outside the hemisphere with their observations) are DECIDED on the device in one call
  // (dba_filter: one byte per observation / point comes back); what follows is the reference's
  // pointer-graph surgery, which leaves the same survivors in the same order.
  std::vector<uint8_t> obs_gone(params_.size(), 0), pt_gone;
  if (!params_.empty()) {
    dba_handle* h = deeparc::engine();
    deeparc::Resident& res = deeparc::resident();
    const bool on_device = deeparc::resident_usable(*this) && !res.pending;
    if (!on_device) {
      // the engine does not hold this scene (first call, or the scene was edited): gather and upload it
      deeparc::resident_invalidate();
      deeparc::flatten(*this, /*freeze_camera=*/false, &res.flat);
      dba_problem view = res.flat.view();
      deeparc::check(dba_problem_set(h, &view), "dba_problem_set");
      res.manager = this;
      res.freeze_camera = 0;
      res.valid = true;
    }
    deeparc::FlatProblem& flat = res.flat;
    pt_gone.assign(flat.point_of.size(), 0);
    deeparc::check(dba_filter(h, error_boundary, hemisphere_center, hemisphere_radius, obs_gone.data(), pt_gone.data(), nullptr, nullptr),
                   "dba_filter");
    for (size_t i = 0; i < params_.size(); ++i)
      if (obs_gone[i]) params_[i]->require_remove(true);  // QUIRK: "<" as written (:348)
    for (size_t i = 0; i < flat.point_of.size(); ++i)
      if (pt_gone[i]) flat.point_of[i]->require_remove(true);
    // the next solve() lets the engine drop them (dba_problem_update) instead of uploading the scene again
    res.obs_remove = obs_gone;
    res.pt_remove = pt_gone;
    res.pending = true;
  } else {
    for (Point3d* p : point3d_) p->require_remove(true);  // nothing observes them: all empty (:368-378)
  }
  params_.erase(std::remove_if(params_.begin(), params_.end(),
                               [](ParameterBlock* b) {
                                 const bool gone = b->require_remove();
                                 if (gone) delete b;  // unlinks itself from its point
                                 return gone;
                               }),
                params_.end());
  point3d_.erase(std::remove_if(point3d_.begin(), point3d_.end(),
                                [](Point3d* p) {
                                  const bool gone = p->require_remove() || p->empty();
                                  if (gone) delete p;
                                  return gone;
                                }),
                 point3d_.end());
}

// `.deeparc` v0.01 text, fixed 6 decimals, points re-indexed, rotations always angle-axis
// (DeepArcManager.cc:426-499).
void DeepArcManager::write(std::string filename) {
  const int64_t n_pts = static_cast<int64_t>(point3d_.size()), n_obs = static_cast<int64_t>(params_.size());
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n_pts; ++i) point3d_[i]->id(static_cast<int>(i));
  std::string head, mid;
  head += "0.010000\n" + std::to_string(params_.size()) + " " + std::to_string(intrinsics_.size()) + " ";
  if (share_extrinsic_)
    head += std::to_string(arc_size_) + " " + std::to_string(ring_size_) + " ";
  else
    head += std::to_string(camera_.size()) + " 0 ";
  head += std::to_string(point3d_.size()) + "\n";
  // the two O(n) sections are formatted by all host cores, one contiguous range per thread
  const int n_thr = std::max(1, omp_get_max_threads());
  std::vector<std::string> obs_part(static_cast<size_t>(n_thr)), pt_part(static_cast<size_t>(n_thr));
#pragma omp parallel num_threads(n_thr)
  {
    const int t = omp_get_thread_num();
    std::string& o = obs_part[t];
    const int64_t b0 = n_obs * t / n_thr, b1 = n_obs * (t + 1) / n_thr;
    o.reserve(static_cast<size_t>(b1 - b0) * 44 + 64);
    for (int64_t i = b0; i < b1; ++i) {
      ParameterBlock* b = params_[i];
      fmti(&o, b->intrinsic()->id());
      o += ' ';
      fmti(&o, share_extrinsic_ ? b->ring()->id() : b->extrinsic()->id());
      o += ' ';
      fmti(&o, b->point3d()->id());
      o += ' ';
      fmt6(&o, b->point2d()->x());
      o += ' ';
      fmt6(&o, b->point2d()->y());
      o += '\n';
    }
    std::string& q = pt_part[t];
    const int64_t p0 = n_pts * t / n_thr, p1 = n_pts * (t + 1) / n_thr;
    q.reserve(static_cast<size_t>(p1 - p0) * 48 + 64);
    for (int64_t i = p0; i < p1; ++i) {
      Point3d* p = point3d_[i];
      for (int j = 0; j < 3; ++j) {
        fmt6(&q, p->position()[j]);
        q += ' ';
      }
      fmti(&q, p->r());
      q += ' ';
      fmti(&q, p->g());
      q += ' ';
      fmti(&q, p->b());
      q += '\n';
    }
  }
  for (Intrinsic* in : intrinsics_) {
    fmt6(&mid, in->center()[0]);
    mid += ' ';
    fmt6(&mid, in->center()[1]);
    mid += ' ' + std::to_string(in->focal_size());
    for (int j = 0; j < in->focal_size(); ++j) {
      mid += ' ';
      fmt6(&mid, in->focal()[j]);
    }
    mid += ' ' + std::to_string(in->distrotion_size());
    for (int j = 0; j < in->distrotion_size(); ++j) {
      mid += ' ';
      fmt6(&mid, in->distrotion()[j]);
    }
    mid += '\n';
  }
  for (Extrinsic* ex : extrinsics_) {
    for (int j = 0; j < 3; ++j) {
      fmt6(&mid, ex->translation()[j]);
      mid += ' ';
    }
    mid += "3 ";
    for (int j = 0; j < 3; ++j) {
      fmt6(&mid, ex->rotation()[j]);
      mid += j < 2 ? " " : "\n";
    }
  }
  std::ofstream of(filename, std::ios::binary);
  of.write(head.data(), static_cast<std::streamsize>(head.size()));
  for (const std::string& s : obs_part) of.write(s.data(), static_cast<std::streamsize>(s.size()));
  of.write(mid.data(), static_cast<std::streamsize>(mid.size()));
  for (const std::string& s : pt_part) of.write(s.data(), static_cast<std::streamsize>(s.size()));
}
