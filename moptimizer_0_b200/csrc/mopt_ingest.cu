// Point-cloud ingest: the data format on the input side of the path.
// The reference reads clouds with `while (file >> x >> y >> z >> r >> g >> b)` (tst/point2point.cpp:125-138,
// fixture tst/data/fachada.txt).  This is a multi-threaded restatement of that loader that parses straight
// into (optionally pinned) host memory ready for mopt_store_upload, plus a raw binary cache format.
// Host-side code: it feeds the device path, it is not a substitute for it.
#include <sys/stat.h>

#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "mopt_internal.h"

using namespace mopt;

namespace {

struct Chunk {
  const char* begin;
  const char* end;
  std::vector<double> vals;  // keep * records
  bool failed = false;       // a malformed / short record ended this chunk early
};

inline const char* skip_ws(const char* p, const char* e) {
  while (p < e && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n' || *p == '\v' || *p == '\f')) ++p;
  return p;
}

// Parses whole records of `columns` numbers; keeps the first `keep` of each.  Stops (failed = true) at the
// first token that is not a number, exactly where the reference's stream extraction would stop.
void parse_chunk(Chunk& c, int columns, int keep) {
  const char* p = c.begin;
  double rec[16];
  for (;;) {
    int got = 0;
    const char* q = p;
    for (; got < columns; ++got) {
      q = skip_ws(q, c.end);
      if (q >= c.end) break;
      if (*q == '+') ++q;  // istream accepts a leading '+', from_chars does not
      double v;
      auto r = std::from_chars(q, c.end, v);
      if (r.ec != std::errc()) {
        c.failed = true;
        return;
      }
      rec[got] = v;
      q = r.ptr;
    }
    if (got < columns) {  // ran out of input: a partial trailing record is dropped
      if (got > 0) c.failed = true;
      return;
    }
    for (int k = 0; k < keep; ++k) c.vals.push_back(rec[k]);
    p = q;
  }
}

struct CloudHeader {
  char magic[8];  // "MOPTCLD1"
  int64_t n;
  int32_t dtype, keep;
};

int alloc_host(void** out, size_t bytes, int pinned) {
  if (bytes == 0) bytes = 16;
  if (pinned) {
    MOPT_CUDA_TRY(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
  } else {
    *out = std::malloc(bytes);
    if (!*out) {
      set_last_error("out of host memory");
      return MOPT_ERR_OUT_OF_MEMORY;
    }
  }
  return MOPT_OK;
}

}  // namespace

extern "C" {

int mopt_cloud_read_text(const char* path, int columns, int keep, int host_dtype, int pinned, void** out, int64_t* n) try {
  MOPT_REQUIRE(path && out && n, "null argument");
  MOPT_REQUIRE(columns >= 1 && columns <= 16 && keep >= 1 && keep <= columns, "bad columns/keep");
  MOPT_REQUIRE(host_dtype == MOPT_F32 || host_dtype == MOPT_F64, "bad host dtype");
  FILE* f = std::fopen(path, "rb");
  if (!f) {
    set_last_error(std::string("not a file! ") + path);  // tst/point2point.cpp:127
    return MOPT_ERR_INVALID_ARGUMENT;
  }
  struct stat sb;
  if (fstat(fileno(f), &sb) != 0 || !S_ISREG(sb.st_mode)) {  // a directory opens fine with "rb" and then lies about its size
    std::fclose(f);
    set_last_error(std::string("not a file! ") + path);
    return MOPT_ERR_INVALID_ARGUMENT;
  }
  const size_t bytes = size_t(sb.st_size);
  std::vector<char> buf;
  try {
    buf.resize(bytes + 1);
  } catch (const std::bad_alloc&) {  // no exception may cross the C ABI
    std::fclose(f);
    set_last_error("out of host memory reading " + std::string(path));
    return MOPT_ERR_OUT_OF_MEMORY;
  }
  const size_t rd = std::fread(buf.data(), 1, bytes, f);
  std::fclose(f);
  buf[rd] = '\n';
  const char* b = buf.data();
  const char* e = b + rd;
  // one chunk per hardware thread, cut at line ends (one record per line, as in the reference's fixture)
  unsigned nt = std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if (rd < (1u << 16)) nt = 1;
  std::vector<Chunk> chunks(nt);
  const char* cur = b;
  for (unsigned t = 0; t < nt; ++t) {
    const char* stop = (t + 1 == nt) ? e : b + size_t(double(rd) * (t + 1) / nt);
    if (stop < cur) stop = cur;
    while (stop < e && *stop != '\n') ++stop;
    chunks[t].begin = cur;
    chunks[t].end = stop;
    cur = stop;
  }
  std::vector<std::thread> th;
  bool oom = false;
  auto guarded = [&oom, columns, keep](Chunk& c) {
    try {
      parse_chunk(c, columns, keep);
    } catch (const std::bad_alloc&) {
      oom = true;  // only ever set, never cleared: a plain bool is enough
    }
  };
  for (unsigned t = 1; t < nt; ++t) th.emplace_back(guarded, std::ref(chunks[t]));
  guarded(chunks[0]);
  for (auto& x : th) x.join();
  if (oom) {
    set_last_error("out of host memory parsing " + std::string(path));
    return MOPT_ERR_OUT_OF_MEMORY;
  }
  size_t total = 0;
  for (unsigned t = 0; t < nt; ++t) {
    total += chunks[t].vals.size();
    if (chunks[t].failed) break;  // everything after the first malformed record is unread, as in the reference
  }
  const size_t esz = host_dtype == MOPT_F32 ? 4 : 8;
  void* dst = nullptr;
  MOPT_TRY(alloc_host(&dst, total * esz, pinned));
  size_t off = 0;
  for (unsigned t = 0; t < nt && off < total; ++t) {
    const std::vector<double>& v = chunks[t].vals;
    if (host_dtype == MOPT_F64) {
      std::memcpy(static_cast<double*>(dst) + off, v.data(), v.size() * 8);
    } else {
      float* o = static_cast<float*>(dst) + off;
      for (size_t i = 0; i < v.size(); ++i) o[i] = float(v[i]);
    }
    off += v.size();
    if (chunks[t].failed) break;
  }
  *out = dst;
  *n = int64_t(total / size_t(keep));
  return MOPT_OK;
}
MOPT_ABI_CATCH

int mopt_cloud_write_binary(const char* path, const void* data, int host_dtype, int keep, int64_t n) try {
  MOPT_REQUIRE(path && (data || n == 0) && n >= 0 && keep >= 1, "bad argument");
  MOPT_REQUIRE(host_dtype == MOPT_F32 || host_dtype == MOPT_F64, "bad host dtype");
  FILE* f = std::fopen(path, "wb");
  if (!f) {
    set_last_error(std::string("cannot open for writing: ") + path);
    return MOPT_ERR_INVALID_ARGUMENT;
  }
  CloudHeader h;
  std::memcpy(h.magic, "MOPTCLD1", 8);
  h.n = n;
  h.dtype = host_dtype;
  h.keep = keep;
  const size_t esz = host_dtype == MOPT_F32 ? 4 : 8;
  bool ok = std::fwrite(&h, sizeof(h), 1, f) == 1;
  ok = ok && (n == 0 || std::fwrite(data, esz * size_t(keep), size_t(n), f) == size_t(n));
  std::fclose(f);
  if (!ok) {
    set_last_error("short write");
    return MOPT_ERR_INVALID_ARGUMENT;
  }
  return MOPT_OK;
}
MOPT_ABI_CATCH

int mopt_cloud_read_binary(const char* path, int pinned, int* host_dtype, int* keep, void** out, int64_t* n) try {
  MOPT_REQUIRE(path && host_dtype && keep && out && n, "null argument");
  FILE* f = std::fopen(path, "rb");
  if (!f) {
    set_last_error(std::string("not a file! ") + path);
    return MOPT_ERR_INVALID_ARGUMENT;
  }
  CloudHeader h;
  if (std::fread(&h, sizeof(h), 1, f) != 1 || std::memcmp(h.magic, "MOPTCLD1", 8) != 0 || h.n < 0 || h.keep < 1 ||
      (h.dtype != MOPT_F32 && h.dtype != MOPT_F64)) {
    std::fclose(f);
    set_last_error("not a MOPTCLD1 cloud file");
    return MOPT_ERR_INVALID_ARGUMENT;
  }
  const size_t esz = h.dtype == MOPT_F32 ? 4 : 8;
  // the header is untrusted: the payload it announces must be exactly what the file holds (this also rules out
  // esz * keep * n wrapping around size_t and under-allocating the buffer fread then fills)
  bool size_ok = std::fseek(f, 0, SEEK_END) == 0;
  const long file_end = size_ok ? std::ftell(f) : -1L;
  size_ok = size_ok && file_end >= long(sizeof(h)) && std::fseek(f, long(sizeof(h)), SEEK_SET) == 0;
  if (size_ok) {
    const unsigned long long payload = (unsigned long long)(file_end) - sizeof(h);
    const unsigned long long rec = (unsigned long long)(esz) * (unsigned long long)(h.keep);
    size_ok = h.keep <= 1024 && (h.n == 0 ? payload == 0 : (payload / rec == (unsigned long long)(h.n) && payload % rec == 0));
  }
  if (!size_ok) {
    std::fclose(f);
    set_last_error("cloud file header does not match the file size (truncated or corrupt)");
    return MOPT_ERR_INVALID_ARGUMENT;
  }
  void* dst = nullptr;
  const int s = alloc_host(&dst, esz * size_t(h.keep) * size_t(h.n), pinned);
  if (s != MOPT_OK) {
    std::fclose(f);
    return s;
  }
  const size_t got = h.n ? std::fread(dst, esz * size_t(h.keep), size_t(h.n), f) : 0;
  std::fclose(f);
  if (got != size_t(h.n)) {
    if (pinned) cudaFreeHost(dst); else std::free(dst);
    set_last_error("truncated cloud file");
    return MOPT_ERR_INVALID_ARGUMENT;
  }
  *host_dtype = h.dtype;
  *keep = h.keep;
  *out = dst;
  *n = h.n;
  return MOPT_OK;
}
MOPT_ABI_CATCH

int mopt_cloud_free(void* p, int pinned) try {
  if (!p) return MOPT_OK;
  if (pinned) MOPT_CUDA_TRY(cudaFreeHost(p)); else std::free(p);
  return MOPT_OK;
}
MOPT_ABI_CATCH

}  // extern "C"
