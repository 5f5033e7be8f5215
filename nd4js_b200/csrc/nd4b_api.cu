// nd4b_api.cu — the C ABI of include/nd4b.h: context (devices, streams, buffers), the host-buffer
// entry points (shard over devices -> chunk -> H2D / kernel / D2H pipelined on per-device streams)
// and the device-resident entry points.  No CPU compute path exists in this library.
#include "../../include/nd4b.h"
#include "kernels.h"
#include <dlfcn.h>
#include <nccl.h>                 // types and prototypes only: the library is dlopen()ed on the first gather, never linked
#include <nvtx3/nvToolsExt.h>   // header-only NVTX 3: ranges cost nothing unless a profiler injects itself

#include <algorithm>
#include <atomic>
#include <climits>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#endif
#include <functional>
#include <map>
#include <unordered_map>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

using nd4b::BatchMap;
using nd4b::BatchMap4;

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

const char* ref_message(int code) {
  switch (code) {
    case ND4B_E_A_NDIM: return "A must be at least 2D.";
    case ND4B_E_B_NDIM: return "B must be at least 2D.";
    case ND4B_E_INNER: return "The last dimension of A and the 2nd to last dimension of B do not match.";
    case ND4B_E_BROADCAST: return "Shapes are not broadcast-compatible.";
    case ND4B_E_NOT_SQUARE: return "Last two dimensions must be quadratic.";
    case ND4B_E_NAN_INPUT: return "Assertion failed.";
    case ND4B_E_SINGULAR: return "Matrix contains NaNs or is (near) singular.";
    default: return "nd4b error";
  }
}

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return fail(ND4B_E_CUDA, "CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, \
                  __LINE__, cudaGetErrorString(e__));                                         \
  } while (0)

constexpr int kSlots = 3;     // pipeline depth per device: chunk c runs on slot c % kSlots
constexpr int kMaxIn = 4, kMaxOut = 3;
constexpr int kOutBase = kMaxIn, kWorkBuf = kMaxIn + kMaxOut;
constexpr int kMaxBuf = kWorkBuf + 1;  // device buffers per slot: up to 3 inputs, 3 outputs, 1 workspace

struct Slot {
  cudaStream_t stream = nullptr;
  cudaEvent_t h2d_done = nullptr;   // recorded after the H2D copies of the chunk in flight: the pinned input ring may be refilled then
  void* buf[kMaxBuf] = {nullptr};
  size_t cap[kMaxBuf] = {0};
  // pinned staging ring for pageable caller memory: [0..2] inputs, [3..5] outputs
  void* pin[kWorkBuf] = {nullptr};
  size_t pin_cap[kWorkBuf] = {0};
  // outputs of the chunk in flight on this slot that still have to be copied from pin[] to the caller
  struct Pending { double* dst; const void* src; size_t bytes; } pending[3];
  int n_pending = 0;
};

// Host-to-host copy with non-temporal stores (SSE2, baseline x86-64).  The staging copies move gigabytes that the CPU does not
// read again (pinned ring -> DMA engine, pinned ring -> caller's result array): ordinary stores first read every destination
// line into the cache (read-for-ownership), i.e. three bytes of DRAM traffic per byte copied, on a host whose memory bandwidth is
// what the pageable path is bound by (the copy threads and the two DMA directions share it); streaming stores make it two.
static void copy_stream(void* dst, const void* src, size_t bytes) {
#if defined(__x86_64__) || defined(_M_X64)
  if (bytes < (size_t)1 << 16) { memcpy(dst, src, bytes); return; }
  char* d = static_cast<char*>(dst);
  const char* s = static_cast<const char*>(src);
  const size_t head = (64 - (reinterpret_cast<uintptr_t>(d) & 63)) & 63;   // up to the first 64-byte line of the destination
  if (head) { memcpy(d, s, head); d += head; s += head; bytes -= head; }
  const size_t lines = bytes / 64;
  for (size_t i = 0; i < lines; i++, d += 64, s += 64) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s));
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 16));
    const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 32));
    const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 48));
    _mm_stream_si128(reinterpret_cast<__m128i*>(d), a);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + 16), b);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + 32), c);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + 48), e);
  }
  _mm_sfence();   // the streamed lines are globally visible before the copy is reported done (the DMA reads them next)
  if (bytes & 63) memcpy(d, s, bytes & 63);
#else
  memcpy(dst, src, bytes);
#endif
}

// A handful of helper threads that split large host-to-host copies (caller memory <-> pinned ring):
// one core moves ~10 GB/s, PCIe Gen5 wants ~55 GB/s.
class CopyPool {
 public:
  explicit CopyPool(int n) {
    for (int i = 0; i < n; i++) workers_.emplace_back([this] { run(); });
  }
  ~CopyPool() {
    { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  void copy(void* dst, const void* src, size_t bytes) {
    const size_t parts = std::min<size_t>(workers_.size() + 1, std::max<size_t>(1, bytes >> 20));
    if (parts <= 1) { copy_stream(dst, src, bytes); return; }
    const size_t step = ((bytes / parts) + 63) & ~size_t(63);
    {
      std::lock_guard<std::mutex> lk(mu_);
      for (size_t i = 1; i < parts; i++) {
        const size_t off = i * step;
        if (off >= bytes) break;
        tasks_.push_back({static_cast<char*>(dst) + off, static_cast<const char*>(src) + off, std::min(step, bytes - off)});
        outstanding_++;
      }
    }
    cv_.notify_all();
    copy_stream(dst, src, std::min(step, bytes));  // the calling thread takes the first part
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return outstanding_ == 0; });
  }

 private:
  struct Task { char* dst; const char* src; size_t bytes; };
  void run() {
    for (;;) {
      Task t;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [this] { return stop_ || !tasks_.empty(); });
        if (stop_ && tasks_.empty()) return;
        t = tasks_.back();
        tasks_.pop_back();
      }
      copy_stream(t.dst, t.src, t.bytes);
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--outstanding_ == 0) done_.notify_all();
      }
    }
  }
  std::vector<std::thread> workers_;
  std::vector<Task> tasks_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  int outstanding_ = 0;
  bool stop_ = false;
};

struct Device {
  int id = -1;
  int sm_count = 0;
  Slot slots[kSlots];
  void* resident[4] = {nullptr, nullptr, nullptr, nullptr};  // whole broadcast operands (matmul, solves: 2; svd_lstsq: 4)
  size_t resident_cap[4] = {0, 0, 0, 0};
  long long* d_info = nullptr;              // cholesky failure key
  int* d_ints = nullptr;                    // [0] = svd sweeps, [1] = svd fail flag
};

struct Context {
  std::vector<Device> devs;
  size_t chunk_bytes = 32u << 20;
  std::mutex mu;
  std::atomic<uint64_t> calls{0}, launches{0}, h2d{0}, d2h{0}, staged{0};
  int last_sweeps = 0;
  CopyPool* pool = nullptr;   // created on first use of pageable memory
  std::vector<ncclComm_t> comms;   // one communicator per device of the context (nd4b_dev_all_gather_f64), created on first use
  double t_copy = 0, t_wait = 0, t_alloc = 0;  // seconds spent in staging copies / stream waits / (re)allocation (ND4B_TRACE)
};

// NVTX range around every host entry point (tracing): `nsys` / `ncu --nvtx` show nd4b_matmul_f64 ... as named ranges.
struct Range {
  explicit Range(const char* name) { nvtxRangePushA(name); }
  ~Range() { nvtxRangePop(); }
};

struct Timer {
  double* acc;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  explicit Timer(double* a) : acc(a) {}
  ~Timer() { *acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

Context* g_ctx = nullptr;
std::mutex g_init_mu;

// cache of idle page-locked blocks behind nd4b_host_alloc / nd4b_host_free
struct PinnedCache {
  std::mutex mu;
  std::map<size_t, std::vector<void*>> idle;   // size class -> blocks
  std::unordered_map<void*, size_t> live;      // blocks handed out -> size class
  size_t idle_bytes = 0;
  size_t limit() const {
    static const size_t v = [] {
      const char* e = getenv("ND4B_PINNED_CACHE_MB");
      return (size_t)(e ? std::max(0L, atol(e)) : 8192L) << 20;
    }();
    return v;
  }
} g_pin;

size_t pinned_class(size_t bytes) {
  if (bytes < 4096) return 4096;
  if (bytes >= (1u << 20)) return (bytes + (1u << 20) - 1) & ~(size_t)((1u << 20) - 1);
  size_t c = 4096;
  while (c < bytes) c <<= 1;
  return c;
}

int ensure(Slot& s, int i, size_t bytes) {
  if (s.cap[i] >= bytes) return 0;
  if (s.buf[i]) CU(cudaFree(s.buf[i]));
  s.buf[i] = nullptr;
  s.cap[i] = 0;
  const size_t want = bytes + bytes / 8 + 256;
  CU(cudaMalloc(&s.buf[i], want));
  s.cap[i] = want;
  return 0;
}

int ensure_pinned(Slot& s, int i, size_t bytes) {
  if (s.pin_cap[i] >= bytes) return 0;
  if (s.pin[i]) CU(cudaFreeHost(s.pin[i]));
  s.pin[i] = nullptr;
  s.pin_cap[i] = 0;
  const size_t want = bytes + bytes / 8 + 256;
  CU(cudaHostAlloc(&s.pin[i], want, cudaHostAllocPortable));
  s.pin_cap[i] = want;
  return 0;
}

// Copies the finished outputs of the chunk that last ran on `slot` from the pinned ring to the caller's memory.
int drain_slot(Context* ctx, Slot& slot) {
  if (slot.n_pending == 0) return 0;
  { Timer t(&ctx->t_wait); CU(cudaStreamSynchronize(slot.stream)); }
  Timer t(&ctx->t_copy);
  for (int i = 0; i < slot.n_pending; i++) ctx->pool->copy(slot.pending[i].dst, slot.pending[i].src, slot.pending[i].bytes);
  slot.n_pending = 0;
  return 0;
}

int ensure_resident(Device& d, int i, size_t bytes) {
  if (d.resident_cap[i] >= bytes) return 0;
  if (d.resident[i]) CU(cudaFree(d.resident[i]));
  d.resident[i] = nullptr;
  d.resident_cap[i] = 0;
  CU(cudaMalloc(&d.resident[i], bytes + 256));
  d.resident_cap[i] = bytes + 256;
  return 0;
}

bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

int init_locked(const int* devices, int n) {
  std::vector<int> ids;
  if (devices && n > 0) ids.assign(devices, devices + n);
  else if (const char* env = getenv("ND4B_DEVICES")) {
    for (const char* p = env; *p;) {
      char* end;
      long v = strtol(p, &end, 10);
      if (end == p) break;
      ids.push_back((int)v);
      p = (*end == ',') ? end + 1 : end;
    }
  }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(ND4B_E_CUDA, "nd4b: no usable CUDA device (%s); this library has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  }
  if (ids.empty()) {
    int cur = 0;
    CU(cudaGetDevice(&cur));
    ids.push_back(cur);
  }
  if (g_ctx) {
    bool same = g_ctx->devs.size() == ids.size();
    for (size_t i = 0; same && i < ids.size(); i++) same = g_ctx->devs[i].id == ids[i];
    if (same) return ND4B_OK;
    return fail(ND4B_E_ARG, "nd4b_init: context already exists with a different device list; call nd4b_shutdown first");
  }
  Context* c = new Context();
  if (const char* mb = getenv("ND4B_CHUNK_MB")) {
    long v = atol(mb);
    if (v > 0) c->chunk_bytes = (size_t)v << 20;
  }
  for (int id : ids) {
    if (id < 0 || id >= count) { delete c; return fail(ND4B_E_ARG, "nd4b_init: device %d out of range (count %d)", id, count); }
    Device d;
    d.id = id;
    CU(cudaSetDevice(id));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, id));
    if (prop.major != 10) { delete c; return fail(ND4B_E_CUDA, "nd4b: device %d is sm_%d%d; this build targets sm_100a (B200) only", id, prop.major, prop.minor); }
    d.sm_count = prop.multiProcessorCount;
    for (auto& s : d.slots) {
      CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
      CU(cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming));
    }
    CU(cudaMalloc(&d.d_info, sizeof(long long)));
    CU(cudaMalloc(&d.d_ints, 4 * sizeof(int)));
    c->devs.push_back(d);
  }
  g_ctx = c;
  return ND4B_OK;
}

int get_ctx(Context** out) {
  std::lock_guard<std::mutex> lk(g_init_mu);
  if (!g_ctx) {
    int rc = init_locked(nullptr, 0);
    if (rc) return rc;
  }
  *out = g_ctx;
  return ND4B_OK;
}

// ---- generic sharded + chunked pipeline --------------------------------------------------------

struct Stream1 {           // one streamed array: `elems` doubles per unit
  const double* in = nullptr;
  double* out = nullptr;
  int64_t elems = 0;
};

struct ChunkArgs {
  Device* dev;
  cudaStream_t stream;
  const double* in[kMaxIn];
  double* out[3];
  double* work;
  size_t work_bytes;
  int64_t count;   // units in this chunk
  int64_t base;    // global index of the first unit
};

// Runs `launch` over `units` independent units: contiguous shards per device, chunks per shard,
// chunk c of a device on slot c % kSlots (H2D -> kernel -> D2H on one stream; slots overlap).
// Chunk sizes ramp up (1/8, 1/4, 1/2 of the nominal chunk, then full chunks) and down again at the end of a shard: the
// pipeline's exposed latency is the H2D of the first chunk plus the D2H of the last one, which at PCIe speed is
// 5-10 % of a whole BASELINE config with 32 MiB chunks.
static int64_t next_chunk_units(int64_t full, int chunk_index, int64_t remaining) {
  const int64_t small = std::max<int64_t>(1, full / 8);
  int64_t cnt = full;
  if (chunk_index < 3) cnt = std::max<int64_t>(small, full >> (3 - chunk_index));
  if (remaining <= 2 * cnt) cnt = std::max<int64_t>(small, (remaining + 1) / 2);
  return std::min(cnt, remaining);
}

// After a failure nothing that was queued may still be running when the call returns: the caller is free to release or
// reuse its buffers, and async copies into them (or kernels on other slots / devices) would write into freed memory.
void quiesce(Context* ctx) {
  for (auto& dv : ctx->devs) {
    cudaSetDevice(dv.id);
    for (auto& sl : dv.slots) {
      cudaStreamSynchronize(sl.stream);
      sl.n_pending = 0;
    }
  }
  cudaGetLastError();
}

int run_pipeline_body(Context* ctx, int64_t units, const std::vector<Stream1>& ins, const std::vector<Stream1>& outs,
                      size_t work_bytes_per_unit, const std::function<int(const ChunkArgs&)>& launch);

int run_pipeline(Context* ctx, int64_t units, const std::vector<Stream1>& ins, const std::vector<Stream1>& outs,
                 size_t work_bytes_per_unit, const std::function<int(const ChunkArgs&)>& launch) {
  const int rc = run_pipeline_body(ctx, units, ins, outs, work_bytes_per_unit, launch);
  if (rc) { const std::string keep = g_err; quiesce(ctx); g_err = keep; }
  return rc;
}

int run_pipeline_body(Context* ctx, int64_t units, const std::vector<Stream1>& ins, const std::vector<Stream1>& outs,
                      size_t work_bytes_per_unit, const std::function<int(const ChunkArgs&)>& launch) {
  const int nd = (int)ctx->devs.size();
  size_t in_bytes_unit = 0, out_bytes_unit = 0;
  for (auto& s : ins) in_bytes_unit += (size_t)s.elems * 8;
  for (auto& s : outs) out_bytes_unit += (size_t)s.elems * 8;
  const size_t per_unit = std::max<size_t>(std::max(in_bytes_unit, out_bytes_unit), 8);
  const int64_t chunk_units = std::max<int64_t>(1, (int64_t)(ctx->chunk_bytes / per_unit));
  std::vector<bool> in_pinned, out_pinned;
  bool any_pageable = false;
  for (auto& s : ins) { in_pinned.push_back(is_pinned(s.in)); any_pageable |= !in_pinned.back(); }
  for (auto& s : outs) { out_pinned.push_back(is_pinned(s.out)); any_pageable |= !out_pinned.back(); }
  if (any_pageable && !ctx->pool) {
    int n = (int)std::thread::hardware_concurrency() / 2 - 1;
    if (const char* e = getenv("ND4B_COPY_THREADS")) n = atoi(e) - 1;
    ctx->pool = new CopyPool(std::max(0, std::min(n, getenv("ND4B_COPY_THREADS") ? 31 : 7)));
  }

  for (auto& dv : ctx->devs)
    for (auto& sl : dv.slots) sl.n_pending = 0;  // nothing may survive from a call that returned early with an error
  struct Shard { int64_t b0, b1, next; int chunk; };
  std::vector<Shard> shards(nd);
  for (int d = 0; d < nd; d++) {
    shards[d].b0 = units * d / nd;
    shards[d].b1 = units * (d + 1) / nd;
    shards[d].next = shards[d].b0;
    shards[d].chunk = 0;
  }
  bool progress = true;
  while (progress) {
    progress = false;
    for (int d = 0; d < nd; d++) {
      Shard& sh = shards[d];
      if (sh.next >= sh.b1) continue;
      progress = true;
      Device& dev = ctx->devs[d];
      CU(cudaSetDevice(dev.id));
      Slot& slot = dev.slots[sh.chunk % kSlots];
      const int64_t cnt = next_chunk_units(chunk_units, sh.chunk, sh.b1 - sh.next);
      // the slot's device buffers are reused in stream order; its pinned ring must be drained first
      if (int rc = drain_slot(ctx, slot)) return rc;
      ChunkArgs a{};
      a.dev = &dev;
      a.stream = slot.stream;
      a.count = cnt;
      a.base = sh.next;
      for (size_t i = 0; i < ins.size(); i++) {
        const size_t bytes = (size_t)cnt * ins[i].elems * 8;
        if (int rc = ensure(slot, (int)i, bytes)) return rc;
        const void* src = ins[i].in + sh.next * ins[i].elems;
        if (!in_pinned[i]) {  // pageable caller memory: stage through the slot's pinned buffer
          { Timer t(&ctx->t_alloc); if (int rc = ensure_pinned(slot, (int)i, bytes)) return rc; }
          // the previous H2D out of this pinned buffer has finished (its kernel and D2H may still be running: an event, not the stream)
          { Timer t(&ctx->t_wait); CU(cudaEventSynchronize(slot.h2d_done)); }
          { Timer t(&ctx->t_copy); ctx->pool->copy(slot.pin[i], src, bytes); }
          src = slot.pin[i];
          ctx->staged += bytes;
        }
        CU(cudaMemcpyAsync(slot.buf[i], src, bytes, cudaMemcpyHostToDevice, slot.stream));
        ctx->h2d += bytes;
        a.in[i] = static_cast<const double*>(slot.buf[i]);
      }
      for (size_t i = 0; i < outs.size(); i++) {
        const size_t bytes = (size_t)cnt * outs[i].elems * 8;
        if (int rc = ensure(slot, kOutBase + (int)i, bytes)) return rc;
        a.out[i] = static_cast<double*>(slot.buf[kOutBase + i]);
      }
      if (any_pageable) CU(cudaEventRecord(slot.h2d_done, slot.stream));
      a.work = nullptr;
      a.work_bytes = 0;
      if (work_bytes_per_unit) {
        a.work_bytes = work_bytes_per_unit * (size_t)cnt;
        if (int rc = ensure(slot, kWorkBuf, a.work_bytes)) return rc;
        a.work = static_cast<double*>(slot.buf[kWorkBuf]);
      }
      if (int rc = launch(a)) return rc;
      for (size_t i = 0; i < outs.size(); i++) {
        const size_t bytes = (size_t)cnt * outs[i].elems * 8;
        double* dst = outs[i].out + sh.next * outs[i].elems;
        if (!out_pinned[i]) {
          if (int rc = ensure_pinned(slot, kOutBase + (int)i, bytes)) return rc;
          CU(cudaMemcpyAsync(slot.pin[kOutBase + i], a.out[i], bytes, cudaMemcpyDeviceToHost, slot.stream));
          slot.pending[slot.n_pending++] = {dst, slot.pin[kOutBase + i], bytes};
          ctx->staged += bytes;
        } else {
          CU(cudaMemcpyAsync(dst, a.out[i], bytes, cudaMemcpyDeviceToHost, slot.stream));
        }
        ctx->d2h += bytes;
      }
      sh.next += cnt;
      sh.chunk++;
    }
  }
  for (int d = 0; d < nd; d++) {
    CU(cudaSetDevice(ctx->devs[d].id));
    for (auto& s : ctx->devs[d].slots) {
      CU(cudaStreamSynchronize(s.stream));
      if (int rc = drain_slot(ctx, s)) return rc;
    }
  }
  if (getenv("ND4B_TRACE"))
    fprintf(stderr, "[nd4b] pipeline: staging copies %.1f ms, stream waits %.1f ms, pinned alloc %.1f ms (cumulative)\n",
            1e3 * ctx->t_copy, 1e3 * ctx->t_wait, 1e3 * ctx->t_alloc);
  return ND4B_OK;
}

// Resets a per-device accumulator (cholesky failure key, svd sweep / failure words) on a slot stream and waits for it: the
// slot streams are non-blocking, so a reset on the legacy stream would have no ordering with the kernels that later
// atomicMin / atomicMax into the word.
int reset_device_words(Device& d, void* dst, const void* src, size_t bytes) {
  CU(cudaSetDevice(d.id));
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, d.slots[0].stream));
  CU(cudaStreamSynchronize(d.slots[0].stream));
  return ND4B_OK;
}

int check_cuda_launch(cudaError_t e, Context* ctx, int n = 1) {
  if (e != cudaSuccess) return fail(ND4B_E_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
  if (ctx) ctx->launches += n;
  return ND4B_OK;
}

// Collapses C's leading dims into the BatchMap odometer (strides in elements, 0 = broadcast).
int build_batch_map(const int32_t* a_shape, int a_ndim, const int32_t* b_shape, int b_ndim,
                    const int32_t* c_shape, int c_ndim, BatchMap* map, bool* a_full, bool* b_full, int64_t* a_count,
                    int64_t* b_count) {
  const int nb = c_ndim - 2;
  const int64_t a_el = (int64_t)a_shape[a_ndim - 2] * a_shape[a_ndim - 1], b_el = (int64_t)b_shape[b_ndim - 2] * b_shape[b_ndim - 1];
  std::vector<int64_t> size(nb), as(nb), bs(nb);
  int64_t sa = a_el, sb = b_el;
  *a_full = true;
  *b_full = true;
  for (int d = nb - 1; d >= 0; d--) {
    const int da = d - c_ndim + a_ndim, db = d - c_ndim + b_ndim;
    const int64_t na = da >= 0 ? a_shape[da] : 1, nbb = db >= 0 ? b_shape[db] : 1;
    size[d] = c_shape[d];
    as[d] = na > 1 ? sa : 0;
    bs[d] = nbb > 1 ? sb : 0;
    if (na != c_shape[d]) *a_full = false;
    if (nbb != c_shape[d]) *b_full = false;
    sa *= na;
    sb *= nbb;
  }
  *a_count = sa / a_el;
  *b_count = sb / b_el;
  // merge adjacent dims when both operands stay affine: stride[d] == size[d+1]*stride[d+1]; drop size-1 dims
  std::vector<int64_t> ms, mas, mbs;
  for (int d = 0; d < nb; d++) {
    if (size[d] == 1) continue;
    if (!ms.empty() && mas.back() == size[d] * as[d] && mbs.back() == size[d] * bs[d]) {
      ms.back() *= size[d];
      mas.back() = as[d];
      mbs.back() = bs[d];
    } else {
      ms.push_back(size[d]);
      mas.push_back(as[d]);
      mbs.push_back(bs[d]);
    }
  }
  if (ms.size() > 8) return fail(ND4B_E_ARG, "matmul: more than 8 non-mergeable broadcast dims are not supported");
  memset(map, 0, sizeof *map);
  map->nd = (int)ms.size();
  for (int d = 0; d < map->nd; d++) { map->size[d] = ms[d]; map->a_str[d] = mas[d]; map->b_str[d] = mbs[d]; }
  return ND4B_OK;
}

// NCCL, loaded at run time for nd4b_dev_all_gather_f64 only
struct NcclApi {
  void* handle = nullptr;
  decltype(&ncclCommInitAll) CommInitAll = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;
} g_nccl;

int load_nccl() {
  if (g_nccl.handle) return ND4B_OK;
  const char* names[] = {getenv("ND4B_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names)
    if (n && (h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
  if (!h) return fail(ND4B_E_CUDA, "nd4b: NCCL is not available (%s)", dlerror());
#define ND4B_NCCL_SYM(field, sym)                                                     \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, #sym));            \
  if (!g_nccl.field) return fail(ND4B_E_CUDA, "nd4b: %s not found in the NCCL library", #sym)
  ND4B_NCCL_SYM(CommInitAll, ncclCommInitAll);
  ND4B_NCCL_SYM(CommDestroy, ncclCommDestroy);
  ND4B_NCCL_SYM(GroupStart, ncclGroupStart);
  ND4B_NCCL_SYM(GroupEnd, ncclGroupEnd);
  ND4B_NCCL_SYM(Broadcast, ncclBroadcast);
  ND4B_NCCL_SYM(GetErrorString, ncclGetErrorString);
  ND4B_NCCL_SYM(GetVersion, ncclGetVersion);
#undef ND4B_NCCL_SYM
  g_nccl.handle = h;
  return ND4B_OK;
}
}  // namespace

// =================================================================================================
extern "C" {

const char* nd4b_last_error(void) { return g_err.c_str(); }
const char* nd4b_version(void) { return "nd4b 0.1 (sm_100a)"; }

int nd4b_init(const int* devices, int n_devices) {
  std::lock_guard<std::mutex> lk(g_init_mu);
  return init_locked(devices, n_devices);
}

int nd4b_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_init_mu);
  if (!g_ctx) return ND4B_OK;
  for (auto& d : g_ctx->devs) {
    cudaSetDevice(d.id);
    for (auto& s : d.slots) {
      if (s.stream) { cudaStreamSynchronize(s.stream); cudaStreamDestroy(s.stream); }
      if (s.h2d_done) cudaEventDestroy(s.h2d_done);
      for (auto& b : s.buf) if (b) cudaFree(b);
      for (auto& b : s.pin) if (b) cudaFreeHost(b);
    }
    for (auto& r : d.resident) if (r) cudaFree(r);
    if (d.d_info) cudaFree(d.d_info);
    if (d.d_ints) cudaFree(d.d_ints);
  }
  if (g_nccl.handle) for (auto c : g_ctx->comms) g_nccl.CommDestroy(c);
  delete g_ctx->pool;
  delete g_ctx;
  g_ctx = nullptr;
  return ND4B_OK;
}

int nd4b_device_count(void) {
  std::lock_guard<std::mutex> lk(g_init_mu);
  return g_ctx ? (int)g_ctx->devs.size() : 0;
}

// Page-locking is expensive (cudaHostAlloc of 512 MiB takes longer than copying it), and every nd.la.* call returns fresh
// arrays: freed blocks are therefore kept in a small cache and handed out again (exact size class: the next multiple of
// 1 MiB for large blocks, the next power of two below that), so that a caller that drops one result and asks for the next
// gets the same pages back.  Bounded by ND4B_PINNED_CACHE_MB (default 8192) of idle blocks; nd4b_host_trim() empties it.
void* nd4b_host_alloc(size_t bytes) {
  const size_t cls = pinned_class(bytes);
  {
    std::lock_guard<std::mutex> lk(g_pin.mu);
    auto it = g_pin.idle.find(cls);
    if (it != g_pin.idle.end() && !it->second.empty()) {
      void* p = it->second.back();
      it->second.pop_back();
      g_pin.idle_bytes -= cls;
      g_pin.live[p] = cls;
      return p;
    }
  }
  void* p = nullptr;
  if (cudaHostAlloc(&p, cls, cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    nd4b_host_trim();   // give the idle blocks back and retry once
    if (cudaHostAlloc(&p, cls, cudaHostAllocPortable) != cudaSuccess) {
      g_err = std::string("nd4b_host_alloc: ") + cudaGetErrorString(cudaGetLastError());
      return nullptr;
    }
  }
  std::lock_guard<std::mutex> lk(g_pin.mu);
  g_pin.live[p] = cls;
  return p;
}

void nd4b_host_free(void* p) {
  if (!p) return;
  size_t cls = 0;
  {
    std::lock_guard<std::mutex> lk(g_pin.mu);
    auto it = g_pin.live.find(p);
    if (it != g_pin.live.end()) {
      cls = it->second;
      g_pin.live.erase(it);
      if (g_pin.idle_bytes + cls <= g_pin.limit()) {
        g_pin.idle[cls].push_back(p);
        g_pin.idle_bytes += cls;
        return;
      }
    }
  }
  cudaFreeHost(p);
}

void nd4b_host_trim(void) {
  std::vector<void*> drop;
  {
    std::lock_guard<std::mutex> lk(g_pin.mu);
    for (auto& kv : g_pin.idle) for (void* q : kv.second) drop.push_back(q);
    g_pin.idle.clear();
    g_pin.idle_bytes = 0;
  }
  for (void* q : drop) cudaFreeHost(q);
}

int nd4b_set_chunk_bytes(size_t bytes) {
  Context* ctx;
  if (int rc = get_ctx(&ctx)) return rc;
  if (bytes < (1u << 16)) return fail(ND4B_E_ARG, "nd4b_set_chunk_bytes: chunk must be at least 64 KiB");
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->chunk_bytes = bytes;
  return ND4B_OK;
}

int nd4b_get_stats(nd4b_stats* out) {
  if (!out) return fail(ND4B_E_ARG, "nd4b_get_stats: null pointer");
  memset(out, 0, sizeof *out);
  std::lock_guard<std::mutex> lk(g_init_mu);
  if (!g_ctx) return ND4B_OK;
  out->calls = g_ctx->calls;
  out->kernel_launches = g_ctx->launches;
  out->h2d_bytes = g_ctx->h2d;
  out->d2h_bytes = g_ctx->d2h;
  out->staged_bytes = g_ctx->staged;
  out->last_sweeps = g_ctx->last_sweeps;
  out->n_devices = (int)g_ctx->devs.size();
  return ND4B_OK;
}

int nd4b_reset_stats(void) {
  std::lock_guard<std::mutex> lk(g_init_mu);
  if (!g_ctx) return ND4B_OK;
  g_ctx->calls = g_ctx->launches = g_ctx->h2d = g_ctx->d2h = g_ctx->staged = 0;
  return ND4B_OK;
}

// ---- matmul -------------------------------------------------------------------------------------

int nd4b_matmul_shape(const int32_t* a_shape, int a_ndim, const int32_t* b_shape, int b_ndim,
                      int32_t* c_shape, int* c_ndim) {
  if (!a_shape || !b_shape || !c_shape || !c_ndim) return fail(ND4B_E_ARG, "matmul_shape: null pointer");
  if (a_ndim < 2) return fail(ND4B_E_A_NDIM, "%s", ref_message(ND4B_E_A_NDIM));
  if (b_ndim < 2) return fail(ND4B_E_B_NDIM, "%s", ref_message(ND4B_E_B_NDIM));
  if (a_ndim > ND4B_MAX_NDIM || b_ndim > ND4B_MAX_NDIM) return fail(ND4B_E_ARG, "matmul: ndim > %d", ND4B_MAX_NDIM);
  for (int d = 0; d < a_ndim; d++) if (a_shape[d] < 1) return fail(ND4B_E_ARG, "Invalid shape: dims must be >= 1.");
  for (int d = 0; d < b_ndim; d++) if (b_shape[d] < 1) return fail(ND4B_E_ARG, "Invalid shape: dims must be >= 1.");
  const int32_t I = a_shape[a_ndim - 2], K = a_shape[a_ndim - 1], J = b_shape[b_ndim - 1];
  if (b_shape[b_ndim - 2] != K) return fail(ND4B_E_INNER, "%s", ref_message(ND4B_E_INNER));
  const int ndim = std::max(a_ndim, b_ndim);
  for (int d = 0; d < ndim; d++) c_shape[d] = 1;
  c_shape[ndim - 2] = I;
  c_shape[ndim - 1] = J;
  const int32_t* shp[2] = {a_shape, b_shape};
  const int nds[2] = {a_ndim, b_ndim};
  for (int w = 0; w < 2; w++)
    for (int i = ndim - 2, j = nds[w] - 2; i-- > 0 && j-- > 0;) {
      if (c_shape[i] == 1) c_shape[i] = shp[w][j];
      else if (c_shape[i] != shp[w][j] && shp[w][j] != 1) return fail(ND4B_E_BROADCAST, "%s", ref_message(ND4B_E_BROADCAST));
    }
  *c_ndim = ndim;
  return ND4B_OK;
}

int nd4b_matmul_f64(const double* A, const int32_t* a_shape, int a_ndim,
                    const double* B, const int32_t* b_shape, int b_ndim,
                    double* C, const int32_t* c_shape, int c_ndim) {
  Range nvtx_range("nd4b_matmul_f64");
  if (!A || !B || !C || !c_shape) return fail(ND4B_E_ARG, "matmul: null pointer");
  int32_t want[ND4B_MAX_NDIM];
  int want_nd = 0;
  if (int rc = nd4b_matmul_shape(a_shape, a_ndim, b_shape, b_ndim, want, &want_nd)) return rc;
  if (want_nd != c_ndim) return fail(ND4B_E_SHAPE, "matmul: result ndim %d, expected %d", c_ndim, want_nd);
  for (int d = 0; d < c_ndim; d++)
    if (want[d] != c_shape[d]) return fail(ND4B_E_SHAPE, "matmul: result shape mismatch at dim %d", d);
  Context* ctx;
  if (int rc = get_ctx(&ctx)) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->calls++;

  const int I = a_shape[a_ndim - 2], K = a_shape[a_ndim - 1], J = b_shape[b_ndim - 1];
  BatchMap map;
  bool a_full, b_full;
  int64_t a_count, b_count;
  if (int rc = build_batch_map(a_shape, a_ndim, b_shape, b_ndim, c_shape, c_ndim, &map, &a_full, &b_full, &a_count, &b_count)) return rc;
  int64_t batch = 1;
  for (int d = 0; d < c_ndim - 2; d++) batch *= c_shape[d];
  const int64_t a_elems = (int64_t)I * K, b_elems = (int64_t)K * J, c_elems = (int64_t)I * J;
  const int nd = (int)ctx->devs.size();

  // Row-panel split of a single large product over the devices: B is replicated, A and C are split by
  // rows.  The rows are the units of the ordinary sharded pipeline (A rows in, C rows out), so that pageable operands go
  // through the pinned ring and the devices work concurrently whatever memory the caller passed.
  if (batch == 1 && nd > 1 && I >= 256 * nd) {
    const size_t bb = (size_t)b_elems * 8;
    for (auto& dev : ctx->devs) {
      CU(cudaSetDevice(dev.id));
      if (int rc = ensure_resident(dev, 1, bb)) return rc;
      CU(cudaMemcpyAsync(dev.resident[1], B, bb, cudaMemcpyHostToDevice, dev.slots[0].stream));
      ctx->h2d += bb;
    }
    for (auto& dev : ctx->devs) {
      CU(cudaSetDevice(dev.id));
      CU(cudaStreamSynchronize(dev.slots[0].stream));
    }
    auto launch_rows = [&](const ChunkArgs& a) -> int {
      BatchMap m1;
      memset(&m1, 0, sizeof m1);
      return check_cuda_launch(nd4b::launch_matmul(a.stream, a.in[0], static_cast<const double*>(a.dev->resident[1]), a.out[0], 1,
                                                   (int)a.count, K, J, m1, a.dev->sm_count), ctx);
    };
    return run_pipeline(ctx, I, {{A, nullptr, (int64_t)K}}, {{nullptr, C, (int64_t)J}}, 0, launch_rows);
  }

  // Operands that follow C's batch index one-to-one are streamed in chunks; operands with any broadcast
  // dim are made resident on every device once and addressed through the odometer.
  std::vector<Stream1> ins, outs;
  int a_slot = -1, b_slot = -1;
  if (a_full) { a_slot = (int)ins.size(); ins.push_back({A, nullptr, a_elems}); }
  if (b_full) { b_slot = (int)ins.size(); ins.push_back({B, nullptr, b_elems}); }
  outs.push_back({nullptr, C, c_elems});
  for (int d = 0; d < nd; d++) {
    Device& dev = ctx->devs[d];
    CU(cudaSetDevice(dev.id));
    if (!a_full) {
      const size_t bytes = (size_t)a_count * a_elems * 8;
      if (int rc = ensure_resident(dev, 0, bytes)) return rc;
      CU(cudaMemcpyAsync(dev.resident[0], A, bytes, cudaMemcpyHostToDevice, dev.slots[0].stream));
      ctx->h2d += bytes;
    }
    if (!b_full) {
      const size_t bytes = (size_t)b_count * b_elems * 8;
      if (int rc = ensure_resident(dev, 1, bytes)) return rc;
      CU(cudaMemcpyAsync(dev.resident[1], B, bytes, cudaMemcpyHostToDevice, dev.slots[0].stream));
      ctx->h2d += bytes;
    }
    if (!a_full || !b_full) CU(cudaStreamSynchronize(dev.slots[0].stream));
  }
  auto launch = [&](const ChunkArgs& a) -> int {
    BatchMap m = map;
    m.base = a.base;
    m.a_lin = a_full ? a_elems : (a_count == 1 ? 0 : -1);
    m.b_lin = b_full ? b_elems : (b_count == 1 ? 0 : -1);
    const double* ap = a_full ? a.in[a_slot] : static_cast<const double*>(a.dev->resident[0]);
    const double* bp = b_full ? a.in[b_slot] : static_cast<const double*>(a.dev->resident[1]);
    return check_cuda_launch(nd4b::launch_matmul(a.stream, ap, bp, a.out[0], a.count, I, K, J, m, a.dev->sm_count), ctx);
  };
  return run_pipeline(ctx, batch, ins, outs, 0, launch);
}

// ---- matmul chain: nd.la.matmul(...matrices), src/la/matmul.js:150-236 ---------------------------
// The parenthesisation (the reference's DP over broadcast-aware flop counts, :159-235) stays with the caller and arrives
// as a postfix plan; here the plan is executed with every intermediate product kept in HBM: operands go up once, one
// result comes down.  Runs on the first device of the context (a chain is a dependent sequence, not a batch to shard).
int nd4b_matmul_plan_f64(int n, const double* const* mats, const int32_t* const* shapes, const int* ndims,
                         const int32_t* plan, int plan_len, double* C, const int32_t* c_shape, int c_ndim) {
  Range nvtx_range("nd4b_matmul_plan_f64");
  if (n < 1 || !mats || !shapes || !ndims || !plan || plan_len < 1 || !C || !c_shape) return fail(ND4B_E_ARG, "matmul_plan: bad argument");
  for (int i = 0; i < n; i++) {
    if (!mats[i] || !shapes[i]) return fail(ND4B_E_ARG, "matmul_plan: null operand %d", i);
    if (ndims[i] < 2) return fail(i == 0 ? ND4B_E_A_NDIM : ND4B_E_B_NDIM, "%s", ref_message(i == 0 ? ND4B_E_A_NDIM : ND4B_E_B_NDIM));
    if (ndims[i] > ND4B_MAX_NDIM) return fail(ND4B_E_ARG, "matmul: ndim > %d", ND4B_MAX_NDIM);
    for (int d = 0; d < ndims[i]; d++) if (shapes[i][d] < 1) return fail(ND4B_E_ARG, "Invalid shape: dims must be >= 1.");
  }
  Context* ctx;
  if (int rc = get_ctx(&ctx)) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->calls++;
  Device& dev = ctx->devs[0];
  CU(cudaSetDevice(dev.id));
  cudaStream_t st = dev.slots[0].stream;

  struct Item { double* p; std::vector<int32_t> shape; };
  std::vector<Item> stack;
  std::vector<void*> owned;
  auto cleanup = [&](int rc) {
    cudaStreamSynchronize(st);
    for (void* q : owned) cudaFree(q);
    return rc;
  };
  auto elems = [](const std::vector<int32_t>& shp) { int64_t e = 1; for (int32_t v : shp) e *= v; return e; };
  for (int t = 0; t < plan_len; t++) {
    const int32_t op = plan[t];
    if (op >= 0) {
      if (op >= n) return cleanup(fail(ND4B_E_ARG, "matmul_plan: operand index %d out of range", op));
      Item it{nullptr, std::vector<int32_t>(shapes[op], shapes[op] + ndims[op])};
      const size_t bytes = (size_t)elems(it.shape) * 8;
      if (cudaMalloc(&it.p, bytes) != cudaSuccess) return cleanup(fail(ND4B_E_CUDA, "matmul_plan: out of device memory"));
      owned.push_back(it.p);
      if (cudaMemcpyAsync(it.p, mats[op], bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) return cleanup(fail(ND4B_E_CUDA, "matmul_plan: H2D copy failed"));
      ctx->h2d += bytes;
      stack.push_back(std::move(it));
      continue;
    }
    if (stack.size() < 2) return cleanup(fail(ND4B_E_ARG, "matmul_plan: malformed plan"));
    Item b = std::move(stack.back()); stack.pop_back();
    Item a = std::move(stack.back()); stack.pop_back();
    const int an = (int)a.shape.size(), bn = (int)b.shape.size();
    Item c{nullptr, std::vector<int32_t>(std::max(an, bn))};
    int cn = 0;
    if (int rc = nd4b_matmul_shape(a.shape.data(), an, b.shape.data(), bn, c.shape.data(), &cn)) return cleanup(rc);
    BatchMap map;
    bool a_full, b_full;
    int64_t a_count, b_count;
    if (int rc = build_batch_map(a.shape.data(), an, b.shape.data(), bn, c.shape.data(), cn, &map, &a_full, &b_full, &a_count, &b_count)) return cleanup(rc);
    const int I = a.shape[an - 2], K = a.shape[an - 1], J = b.shape[bn - 1];
    int64_t batch = 1;
    for (int d = 0; d < cn - 2; d++) batch *= c.shape[d];
    map.base = 0;
    map.a_lin = a_full ? (int64_t)I * K : (a_count == 1 ? 0 : -1);
    map.b_lin = b_full ? (int64_t)K * J : (b_count == 1 ? 0 : -1);
    if (cudaMalloc(&c.p, (size_t)elems(c.shape) * 8) != cudaSuccess) return cleanup(fail(ND4B_E_CUDA, "matmul_plan: out of device memory"));
    owned.push_back(c.p);
    if (int rc = check_cuda_launch(nd4b::launch_matmul(st, a.p, b.p, c.p, batch, I, K, J, map, dev.sm_count), ctx)) return cleanup(rc);
    stack.push_back(std::move(c));
  }
  if (stack.size() != 1) return cleanup(fail(ND4B_E_ARG, "matmul_plan: malformed plan"));
  const Item& r = stack.back();
  if ((int)r.shape.size() != c_ndim) return cleanup(fail(ND4B_E_SHAPE, "matmul: result ndim %d, expected %d", c_ndim, (int)r.shape.size()));
  for (int d = 0; d < c_ndim; d++)
    if (r.shape[d] != c_shape[d]) return cleanup(fail(ND4B_E_SHAPE, "matmul: result shape mismatch at dim %d", d));
  const size_t bytes = (size_t)elems(r.shape) * 8;
  if (cudaMemcpyAsync(C, r.p, bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess) return cleanup(fail(ND4B_E_CUDA, "matmul_plan: D2H copy failed"));
  ctx->d2h += bytes;
  return cleanup(ND4B_OK);
}

// ---- cholesky -----------------------------------------------------------------------------------

int nd4b_cholesky_f64(const double* S, double* L, int64_t batch, int n, int64_t* first_bad) {
  Range nvtx_range("nd4b_cholesky_f64");
  if (first_bad) *first_bad = -1;
  if (!S || !L) return fail(ND4B_E_ARG, "cholesky: null pointer");
  if (batch < 1 || n < 1) return fail(ND4B_E_ARG, "cholesky: batch and n must be >= 1");
  Context* ctx;
  if (int rc = get_ctx(&ctx)) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->calls++;
  const long long none = LLONG_MAX;
  for (auto& d : ctx->devs)
    if (int rc = reset_device_words(d, d.d_info, &none, sizeof none)) return rc;
  const int64_t nn = (int64_t)n * n;
  auto launch = [&](const ChunkArgs& a) -> int {
    return check_cuda_launch(nd4b::launch_cholesky(a.stream, a.in[0], a.out[0], a.count, n, a.dev->d_info, a.base), ctx);
  };
  if (int rc = run_pipeline(ctx, batch, {{S, nullptr, nn}}, {{nullptr, L, nn}}, 0, launch)) return rc;
  long long key = LLONG_MAX;
  for (auto& d : ctx->devs) {
    long long k;
    CU(cudaSetDevice(d.id));
    CU(cudaMemcpy(&k, d.d_info, sizeof k, cudaMemcpyDeviceToHost));
    key = std::min(key, k);
  }
  if (key != LLONG_MAX) {
    if (first_bad) *first_bad = key >> 1;
    const int code = (key & 1) ? ND4B_E_SINGULAR : ND4B_E_NAN_INPUT;
    return fail(code, "%s", ref_message(code));
  }
  return ND4B_OK;
}

// ---- qr -----------------------------------------------------------------------------------------

int nd4b_qr_f64(const double* A, double* Q, double* R, int64_t batch, int rows, int cols) {
  Range nvtx_range("nd4b_qr_f64");
  if (!A || !Q || !R) return fail(ND4B_E_ARG, "qr: null pointer");
  if (batch < 1 || rows < 1 || cols < 1) return fail(ND4B_E_ARG, "qr: batch, rows and cols must be >= 1");
  Context* ctx;
  if (int rc = get_ctx(&ctx)) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->calls++;
  const int L = std::min(rows, cols);
  const size_t work_unit = nd4b::qr_workspace_bytes(1, rows, cols);
  auto launch = [&](const ChunkArgs& a) -> int {
    return check_cuda_launch(nd4b::launch_qr(a.stream, a.in[0], a.out[0], a.out[1], a.count, rows, cols, a.work, a.work_bytes), ctx);
  };
  return run_pipeline(ctx, batch, {{A, nullptr, (int64_t)rows * cols}},
                      {{nullptr, Q, (int64_t)rows * L}, {nullptr, R, (int64_t)L * cols}}, work_unit, launch);
}

int nd4b_qr_inplace_f64(const double* A, const double* Y, double* R, double* QtY, int64_t batch, int M, int N, int L) {
  Range nvtx_range("nd4b_qr_inplace_f64");
  if (!A || !Y || !R || !QtY) return fail(ND4B_E_ARG, "qr_inplace: null pointer");
  if (batch < 1 || M < 1 || N < 1 || L < 1) return fail(ND4B_E_ARG, "qr_inplace: batch, M, N and L must be >= 1");
  Context* ctx;
  if (int rc = get_ctx(&ctx)) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->calls++;
  auto launch = [&](const ChunkArgs& a) -> int {
    return check_cuda_launch(nd4b::launch_qr_inplace(a.stream, a.in[0], a.in[1], a.out[0], a.out[1], a.count, M, N, L), ctx);
  };
  return run_pipeline(ctx, batch, {{A, nullptr, (int64_t)M * N}, {Y, nullptr, (int64_t)M * L}},
                      {{nullptr, R, (int64_t)M * N}, {nullptr, QtY, (int64_t)M * L}}, 0, launch);
}

// ---- qr_lstsq (src/la/qr.js:186-273), fused form for thin factors ------------------------------

int nd4b_qr_lstsq_f64(const double* Q, const double* R, const double* Y, double* X, int64_t batch, int N, int M, int I, int J) {
  Range nvtx_range("nd4b_qr_lstsq_f64");
  if (!Q || !R || !Y || !X) return fail(ND4B_E_ARG, "qr_lstsq: null pointer");
  if (batch < 1 || N < 1 || M < 1 || I < 1 || J < 1) return fail(ND4B_E_ARG, "qr_lstsq: batch, N, M, I and J must be >= 1");
  if (I > N) return fail(ND4B_E_ARG, "qr_lstsq(Q,R,y): Under-determined systems not supported. Use rrqr instead.");
  if (M > 32 || I > 32) return fail(ND4B_E_ARG, "qr_lstsq: the fused kernel takes factors with at most 32 columns; compose matmul2 and triu_solve");
  Context* ctx;
  if (int rc = get_ctx(&ctx)) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->calls++;
  auto launch = [&](const ChunkArgs& a) -> int {
    return check_cuda_launch(nd4b::launch_qr_lstsq(a.stream, a.in[0], a.in[1], a.in[2], a.out[0], a.count, N, M, I, J), ctx);
  };
  return run_pipeline(ctx, batch, {{Q, nullptr, (int64_t)N * M}, {R, nullptr, (int64_t)M * I}, {Y, nullptr, (int64_t)N * J}},
                      {{nullptr, X, (int64_t)I * J}}, 0, launch);
}

// ---- svd ----------------------------------------------------------------------------------------

int nd4b_svd_jac1_f64(const double* A, double* U, double* sv, double* V,
                      int64_t batch, int rows, int cols, int* sweeps_out) {
  Range nvtx_range("nd4b_svd_jac1_f64");
  if (sweeps_out) *sweeps_out = 0;
  if (!A || !U || !sv || !V) return fail(ND4B_E_ARG, "svd_jac_1sided: null pointer");
  if (batch < 1 || rows < 1 || cols < 1) return fail(ND4B_E_ARG, "svd_jac_1sided: batch, rows and cols must be >= 1");
  Context* ctx;
  if (int rc = get_ctx(&ctx)) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->calls++;
  const int zeros[4] = {0, 0, 0, 0};
  for (auto& d : ctx->devs)
    if (int rc = reset_device_words(d, d.d_ints, zeros, sizeof zeros)) return rc;
  const int L = std::min(rows, cols);
  const size_t work_unit = nd4b::svd_workspace_bytes(1, rows, cols);
  auto launch = [&](const ChunkArgs& a) -> int {
    const int n_kernels = (rows == 64 && cols == 64 && a.work_bytes) ? 3 : 1;   // preconditioner: FP32 Jacobi, V1 / G1, FP64 Jacobi
    return check_cuda_launch(nd4b::launch_svd_jac1(a.stream, a.in[0], a.out[0], a.out[1], a.out[2], a.count, rows, cols,
                                                   a.dev->d_ints, a.dev->d_ints + 1, a.work, a.work_bytes), ctx, n_kernels);
  };
  if (int rc = run_pipeline(ctx, batch, {{A, nullptr, (int64_t)rows * cols}},
                            {{nullptr, U, (int64_t)rows * L}, {nullptr, sv, (int64_t)L}, {nullptr, V, (int64_t)L * cols}},
                            work_unit, launch)) return rc;
  int sweeps = 0, failed = 0;
  for (auto& d : ctx->devs) {
    int h[2];
    CU(cudaSetDevice(d.id));
    CU(cudaMemcpy(h, d.d_ints, sizeof h, cudaMemcpyDeviceToHost));
    sweeps = std::max(sweeps, h[0]);
    failed |= h[1];
  }
  ctx->last_sweeps = sweeps;
  if (sweeps_out) *sweeps_out = sweeps;
  if (failed) return fail(ND4B_E_NO_CONVERGENCE, "svd_jac_1sided: no convergence within the sweep limit");
  return ND4B_OK;
}

// ---- svd_rank / svd_lstsq / svd_solve (src/la/svd.js:31-226) -------------------------------------

int nd4b_svd_rank_f64(const double* sv, int32_t* rank, int64_t batch, int n) {
  Range nvtx_range("nd4b_svd_rank_f64");
  if (!sv || !rank) return fail(ND4B_E_ARG, "svd_rank: null pointer");
  if (batch < 1 || n < 1) return fail(ND4B_E_ARG, "svd_rank: batch and n must be >= 1");
  Context* ctx;
  if (int rc = get_ctx(&ctx)) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->calls++;
  // O(batch * n) integer work: one device, one chunk (the call exists so that the rank rule lives behind the boundary)
  Device& dev = ctx->devs[0];
  Slot& slot = dev.slots[0];
  CU(cudaSetDevice(dev.id));
  const size_t in_bytes = (size_t)batch * n * 8, out_bytes = (size_t)batch * 4;
  if (int rc = ensure(slot, 0, in_bytes)) return rc;
  if (int rc = ensure(slot, kOutBase, out_bytes)) return rc;
  const int zeros[4] = {0, 0, 0, 0};
  if (int rc = reset_device_words(dev, dev.d_ints, zeros, sizeof zeros)) return rc;
  CU(cudaMemcpyAsync(slot.buf[0], sv, in_bytes, cudaMemcpyHostToDevice, slot.stream));
  if (int rc = check_cuda_launch(nd4b::launch_svd_rank(slot.stream, (const double*)slot.buf[0], (int*)slot.buf[kOutBase], batch, n,
                                                       dev.d_ints + 2), ctx)) { quiesce(ctx); return rc; }
  int bad = 0;
  cudaError_t e1 = cudaMemcpyAsync(rank, slot.buf[kOutBase], out_bytes, cudaMemcpyDeviceToHost, slot.stream);
  cudaError_t e2 = cudaMemcpyAsync(&bad, dev.d_ints + 2, sizeof bad, cudaMemcpyDeviceToHost, slot.stream);
  cudaError_t e3 = cudaStreamSynchronize(slot.stream);
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) { quiesce(ctx); return fail(ND4B_E_CUDA, "svd_rank: copy failed"); }
  ctx->h2d += in_bytes;
  ctx->d2h += out_bytes;
  if (bad) return fail(ND4B_E_NAN_INPUT, "svd_rank(): NaN or Infinity encountered.");
  return ND4B_OK;
}

namespace {
struct Operand4 { const int32_t* shape; int ndim; int lead; int64_t elems; const char* name; };

// Validation and broadcast shape of svd_lstsq(U,sv,V,y) with the reference's checks and texts (svd.js:112-147).
int svd_lstsq_shape_impl(const Operand4 (&op)[4], int32_t* x_shape, int* x_ndim) {
  static const char* nd_msg[4] = {"svd_lstsq(U,sv,V, y): U.ndim must be at least 2.", "svd_lstsq(U,sv,V, y): sv.ndim must be at least 1.",
                                  "svd_lstsq(U,sv,V, y): V.ndim must be at least 2.", "svd_lstsq(U,sv,V, y): y.ndim must be at least 2."};
  for (int o = 0; o < 4; o++) {
    if (!op[o].shape) return fail(ND4B_E_ARG, "svd_lstsq: null shape");
    if (op[o].ndim < (o == 1 ? 1 : 2)) return fail(o == 3 ? ND4B_E_B_NDIM : ND4B_E_A_NDIM, "%s", nd_msg[o]);
    if (op[o].ndim > ND4B_MAX_NDIM - 1) return fail(ND4B_E_ARG, "svd_lstsq: ndim > %d", ND4B_MAX_NDIM - 1);
    for (int d = 0; d < op[o].ndim; d++) if (op[o].shape[d] < 1) return fail(ND4B_E_ARG, "Invalid shape: dims must be >= 1.");
  }
  const int32_t *us = op[0].shape, *ss = op[1].shape, *vs = op[2].shape, *ys = op[3].shape;
  const int N = us[op[0].ndim - 2], M = us[op[0].ndim - 1], I = vs[op[2].ndim - 1], J = ys[op[3].ndim - 1];
  if (N != ys[op[3].ndim - 2]) return fail(ND4B_E_INNER, "svd_lstsq(U,sv,V, y): U and y don't match.");
  if (M != ss[op[1].ndim - 1]) return fail(ND4B_E_INNER, "svd_lstsq(U,sv,V, y): U and sv don't match.");
  if (M != vs[op[2].ndim - 2]) return fail(ND4B_E_INNER, "svd_lstsq(U,sv,V, y): V and sv don't match.");
  const int ndim = std::max(std::max(op[0].ndim, op[1].ndim + 1), std::max(op[2].ndim, op[3].ndim));
  for (int d = 0; d < ndim; d++) x_shape[d] = 1;
  x_shape[ndim - 2] = I;
  x_shape[ndim - 1] = J;
  static const int order[4] = {0, 2, 3, 1};   // U, V, y, then sv (svd.js:132-145)
  for (int w = 0; w < 4; w++) {
    const Operand4& a = op[order[w]];
    for (int i = ndim - 2, j = a.lead; i-- > 0 && j-- > 0;) {
      if (x_shape[i] == 1) x_shape[i] = a.shape[j];
      else if (x_shape[i] != a.shape[j] && a.shape[j] != 1)
        return fail(ND4B_E_BROADCAST, "svd_lstsq(U,sv,V, y): U,sv,V,y not broadcast-compatible.");
    }
  }
  *x_ndim = ndim;
  return ND4B_OK;
}
}  // namespace

int nd4b_svd_lstsq_shape(const int32_t* u_shape, int u_ndim, const int32_t* sv_shape, int sv_ndim,
                         const int32_t* v_shape, int v_ndim, const int32_t* y_shape, int y_ndim,
                         int32_t* x_shape, int* x_ndim) {
  if (!x_shape || !x_ndim) return fail(ND4B_E_ARG, "svd_lstsq_shape: null pointer");
  const Operand4 op[4] = {{u_shape, u_ndim, u_ndim - 2, 0, "U"}, {sv_shape, sv_ndim, sv_ndim - 1, 0, "sv"},
                          {v_shape, v_ndim, v_ndim - 2, 0, "V"}, {y_shape, y_ndim, y_ndim - 2, 0, "y"}};
  return svd_lstsq_shape_impl(op, x_shape, x_ndim);
}

int nd4b_svd_lstsq_f64(const double* U, const int32_t* u_shape, int u_ndim, const double* sv, const int32_t* sv_shape, int sv_ndim,
                       const double* V, const int32_t* v_shape, int v_ndim, const double* Y, const int32_t* y_shape, int y_ndim,
                       double* X, const int32_t* x_shape, int x_ndim) {
  Range nvtx_range("nd4b_svd_lstsq_f64");
  if (!U || !sv || !V || !Y || !X || !x_shape) return fail(ND4B_E_ARG, "svd_lstsq: null pointer");
  Operand4 op[4] = {{u_shape, u_ndim, u_ndim - 2, 0, "U"}, {sv_shape, sv_ndim, sv_ndim - 1, 0, "sv"},
                    {v_shape, v_ndim, v_ndim - 2, 0, "V"}, {y_shape, y_ndim, y_ndim - 2, 0, "y"}};
  int32_t want[ND4B_MAX_NDIM];
  int ndim = 0;
  if (int rc = svd_lstsq_shape_impl(op, want, &ndim)) return rc;
  if (ndim != x_ndim) return fail(ND4B_E_SHAPE, "svd_lstsq: result ndim %d, expected %d", x_ndim, ndim);
  for (int d = 0; d < ndim; d++)
    if (want[d] != x_shape[d]) return fail(ND4B_E_SHAPE, "svd_lstsq: result shape mismatch at dim %d", d);
  const int N = u_shape[u_ndim - 2], M = u_shape[u_ndim - 1], I = v_shape[v_ndim - 1], J = y_shape[y_ndim - 1];
  if ((size_t)M * J * 8 > 200 * 1024) return fail(ND4B_E_ARG, "svd_lstsq: M*J = %d*%d exceeds the shared-memory tile of the kernel", M, J);
  op[0].elems = (int64_t)N * M; op[1].elems = M; op[2].elems = (int64_t)M * I; op[3].elems = (int64_t)N * J;
  const double* ptr[4] = {U, sv, V, Y};

  Context* ctx;
  if (int rc = get_ctx(&ctx)) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->calls++;

  // odometer over the result's leading dims: strides per operand (0 = broadcast), adjacent dims merged when all four stay affine
  const int nb = ndim - 2;
  std::vector<int64_t> size(nb);
  std::vector<int64_t> str[4];
  bool full[4];
  int64_t count[4];
  for (int o = 0; o < 4; o++) {
    str[o].assign(nb, 0);
    full[o] = true;
    int64_t s = op[o].elems;
    for (int d = nb - 1; d >= 0; d--) {
      const int idx = d - nb + op[o].lead;
      const int64_t n = idx >= 0 ? op[o].shape[idx] : 1;
      str[o][d] = n > 1 ? s : 0;
      if (n != x_shape[d]) full[o] = false;
      s *= n;
    }
    count[o] = s / op[o].elems;
  }
  BatchMap4 map;
  memset(&map, 0, sizeof map);
  int64_t batch = 1;
  {
    std::vector<int64_t> ms;
    std::vector<int64_t> mstr[4];
    for (int d = 0; d < nb; d++) {
      size[d] = x_shape[d];
      batch *= size[d];
      if (size[d] == 1) continue;
      bool merge = !ms.empty();
      for (int o = 0; o < 4 && merge; o++) merge = mstr[o].back() == size[d] * str[o][d];
      if (merge) {
        ms.back() *= size[d];
        for (int o = 0; o < 4; o++) mstr[o].back() = str[o][d];
      } else {
        ms.push_back(size[d]);
        for (int o = 0; o < 4; o++) mstr[o].push_back(str[o][d]);
      }
    }
    if (ms.size() > 8) return fail(ND4B_E_ARG, "svd_lstsq: more than 8 non-mergeable broadcast dims are not supported");
    map.nd = (int)ms.size();
    for (int d = 0; d < map.nd; d++) {
      map.size[d] = ms[d];
      for (int o = 0; o < 4; o++) map.str[o][d] = mstr[o][d];
    }
  }
  const int zeros[4] = {0, 0, 0, 0};
  for (auto& d : ctx->devs)
    if (int rc = reset_device_words(d, d.d_ints, zeros, sizeof zeros)) return rc;
  std::vector<Stream1> ins, outs;
  int slot_of[4] = {-1, -1, -1, -1};
  for (int o = 0; o < 4; o++)
    if (full[o]) { slot_of[o] = (int)ins.size(); ins.push_back({ptr[o], nullptr, op[o].elems}); }
  outs.push_back({nullptr, X, (int64_t)I * J});
  for (auto& dev : ctx->devs) {
    CU(cudaSetDevice(dev.id));
    bool any = false;
    for (int o = 0; o < 4; o++) {
      if (full[o]) continue;
      const size_t bytes = (size_t)count[o] * op[o].elems * 8;
      if (int rc = ensure_resident(dev, o, bytes)) return rc;
      CU(cudaMemcpyAsync(dev.resident[o], ptr[o], bytes, cudaMemcpyHostToDevice, dev.slots[0].stream));
      ctx->h2d += bytes;
      any = true;
    }
    if (any) CU(cudaStreamSynchronize(dev.slots[0].stream));
  }
  auto launch = [&](const ChunkArgs& a) -> int {
    BatchMap4 m = map;
    m.base = a.base;
    const double* p[4];
    for (int o = 0; o < 4; o++) {
      m.lin[o] = full[o] ? op[o].elems : (count[o] == 1 ? 0 : -1);
      p[o] = full[o] ? a.in[slot_of[o]] : static_cast<const double*>(a.dev->resident[o]);
    }
    return check_cuda_launch(nd4b::launch_svd_lstsq(a.stream, p[0], p[1], p[2], p[3], a.out[0], a.count, N, M, I, J, m, a.dev->d_ints + 2), ctx);
  };
  if (int rc = run_pipeline(ctx, batch, ins, outs, 0, launch)) return rc;
  int bad = 0;
  for (auto& d : ctx->devs) {
    int h = 0;
    CU(cudaSetDevice(d.id));
    CU(cudaMemcpy(&h, d.d_ints + 2, sizeof h, cudaMemcpyDeviceToHost));
    bad |= h;
  }
  if (bad) return fail(ND4B_E_NAN_INPUT, "svd_solve(): NaN or Infinity encountered.");
  return ND4B_OK;
}

// ---- triangular solves ----------------------------------------------------------------------------

int nd4b_tri_solve_f64(int op, const double* T, const int32_t* t_shape, int t_ndim,
                       const double* Y, const int32_t* y_shape, int y_ndim,
                       double* X, const int32_t* x_shape, int x_ndim) {
  Range nvtx_range("nd4b_tri_solve_f64");
  static const char* who[3] = {"tril_solve(L,Y)", "triu_solve(U,Y)", "cholesky_solve(L,y)"};
  static const char* tn[3] = {"L", "U", "L"};
  if (op < 0 || op > 2) return fail(ND4B_E_ARG, "tri_solve: op must be 0, 1 or 2");
  if (!T || !Y || !X || !t_shape || !y_shape || !x_shape) return fail(ND4B_E_ARG, "%s: null pointer", who[op]);
  if (t_ndim < 2) return fail(ND4B_E_A_NDIM, op == 2 ? "L must be at least 2D." : "%s: %s.ndim must be at least 2.", who[op], tn[op]);
  if (y_ndim < 2) return fail(ND4B_E_B_NDIM, op == 2 ? "y must be at least 2D." : "%s: Y.ndim must be at least 2.", who[op]);
  if (t_ndim > ND4B_MAX_NDIM || y_ndim > ND4B_MAX_NDIM) return fail(ND4B_E_ARG, "%s: ndim > %d", who[op], ND4B_MAX_NDIM);
  for (int d = 0; d < t_ndim; d++) if (t_shape[d] < 1) return fail(ND4B_E_ARG, "Invalid shape: dims must be >= 1.");
  for (int d = 0; d < y_ndim; d++) if (y_shape[d] < 1) return fail(ND4B_E_ARG, "Invalid shape: dims must be >= 1.");
  const int M = y_shape[y_ndim - 2], J = y_shape[y_ndim - 1];
  if (op == 2) {  // cholesky.js:84-85 checks squareness first
    if (t_shape[t_ndim - 2] != t_shape[t_ndim - 1]) return fail(ND4B_E_NOT_SQUARE, "Last two dimensions of L must be quadratic.");
    if (t_shape[t_ndim - 1] != M) return fail(ND4B_E_INNER, "L and y don't match.");
  } else {
    if (t_shape[t_ndim - 2] != M) return fail(ND4B_E_INNER, "%s: %s and Y don't match.", who[op], tn[op]);
    if (t_shape[t_ndim - 1] != M) return fail(ND4B_E_NOT_SQUARE, "%s: Last two dimensions of %s must be quadratic.", who[op], tn[op]);
  }
  const int ndim = std::max(t_ndim, y_ndim);
  int32_t want[ND4B_MAX_NDIM];
  for (int d = 0; d < ndim; d++) want[d] = 1;
  want[ndim - 2] = M;
  want[ndim - 1] = J;
  const int32_t* shp[2] = {t_shape, y_shape};
  const int nds[2] = {t_ndim, y_ndim};
  for (int w = 0; w < 2; w++)
    for (int i = ndim - 2, j = nds[w] - 2; i-- > 0 && j-- > 0;) {
      if (want[i] == 1) want[i] = shp[w][j];
      else if (want[i] != shp[w][j] && shp[w][j] != 1)
        return fail(ND4B_E_BROADCAST, op == 2 ? "Shapes are not broadcast-compatible." : "%s: %s and Y not broadcast-compatible.", who[op], tn[op]);
    }
  if (x_ndim != ndim) return fail(ND4B_E_SHAPE, "%s: result ndim %d, expected %d", who[op], x_ndim, ndim);
  for (int d = 0; d < ndim; d++)
    if (want[d] != x_shape[d]) return fail(ND4B_E_SHAPE, "%s: result shape mismatch at dim %d", who[op], d);

  Context* ctx;
  if (int rc = get_ctx(&ctx)) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->calls++;
  BatchMap map;
  bool t_full, y_full;
  int64_t t_count, y_count;
  if (int rc = build_batch_map(t_shape, t_ndim, y_shape, y_ndim, x_shape, ndim, &map, &t_full, &y_full, &t_count, &y_count)) return rc;
  int64_t batch = 1;
  for (int d = 0; d < ndim - 2; d++) batch *= x_shape[d];
  const int64_t t_elems = (int64_t)M * M, y_elems = (int64_t)M * J;
  std::vector<Stream1> ins, outs;
  int t_slot = -1, y_slot = -1;
  if (t_full) { t_slot = (int)ins.size(); ins.push_back({T, nullptr, t_elems}); }
  if (y_full) { y_slot = (int)ins.size(); ins.push_back({Y, nullptr, y_elems}); }
  outs.push_back({nullptr, X, y_elems});
  for (auto& dev : ctx->devs) {
    CU(cudaSetDevice(dev.id));
    if (!t_full) {
      const size_t bytes = (size_t)t_count * t_elems * 8;
      if (int rc = ensure_resident(dev, 0, bytes)) return rc;
      CU(cudaMemcpyAsync(dev.resident[0], T, bytes, cudaMemcpyHostToDevice, dev.slots[0].stream));
      ctx->h2d += bytes;
    }
    if (!y_full) {
      const size_t bytes = (size_t)y_count * y_elems * 8;
      if (int rc = ensure_resident(dev, 1, bytes)) return rc;
      CU(cudaMemcpyAsync(dev.resident[1], Y, bytes, cudaMemcpyHostToDevice, dev.slots[0].stream));
      ctx->h2d += bytes;
    }
    if (!t_full || !y_full) CU(cudaStreamSynchronize(dev.slots[0].stream));
  }
  auto launch = [&](const ChunkArgs& a) -> int {
    BatchMap mm = map;
    mm.base = a.base;
    mm.a_lin = t_full ? t_elems : (t_count == 1 ? 0 : -1);
    mm.b_lin = y_full ? y_elems : (y_count == 1 ? 0 : -1);
    const double* tp = t_full ? a.in[t_slot] : static_cast<const double*>(a.dev->resident[0]);
    const double* yp = y_full ? a.in[y_slot] : static_cast<const double*>(a.dev->resident[1]);
    return check_cuda_launch(nd4b::launch_tri_solve(a.stream, op, tp, yp, a.out[0], a.count, M, J, mm), ctx);
  };
  return run_pipeline(ctx, batch, ins, outs, 0, launch);
}

// ---- device-resident forms ----------------------------------------------------------------------

static int dev_enter(int device, Context** ctx, int* sm_count) {
  if (int rc = get_ctx(ctx)) return rc;
  CU(cudaSetDevice(device));
  *sm_count = 148;
  for (auto& d : (*ctx)->devs) if (d.id == device) *sm_count = d.sm_count;
  return ND4B_OK;
}

int nd4b_dev_matmul_f64(int device, void* stream, const double* A, int64_t a_stride,
                        const double* B, int64_t b_stride, double* C, int64_t batch, int I, int K, int J) {
  if (!A || !B || !C || batch < 1 || I < 1 || K < 1 || J < 1 || a_stride < 0 || b_stride < 0)
    return fail(ND4B_E_ARG, "dev_matmul: bad argument");
  Context* ctx; int sms;
  if (int rc = dev_enter(device, &ctx, &sms)) return rc;
  BatchMap m;
  memset(&m, 0, sizeof m);
  m.a_lin = a_stride;
  m.b_lin = b_stride;
  return check_cuda_launch(nd4b::launch_matmul((cudaStream_t)stream, A, B, C, batch, I, K, J, m, sms), ctx);
}

int nd4b_dev_cholesky_f64(int device, void* stream, const double* S, double* L, int64_t batch, int n, long long* info) {
  if (!S || !L || batch < 1 || n < 1) return fail(ND4B_E_ARG, "dev_cholesky: bad argument");
  Context* ctx; int sms;
  if (int rc = dev_enter(device, &ctx, &sms)) return rc;
  return check_cuda_launch(nd4b::launch_cholesky((cudaStream_t)stream, S, L, batch, n, info, 0), ctx);
}

size_t nd4b_dev_qr_workspace(int64_t batch, int rows, int cols) { return nd4b::qr_workspace_bytes(batch, rows, cols); }
size_t nd4b_dev_svd_workspace(int64_t batch, int rows, int cols) { return nd4b::svd_workspace_bytes(batch, rows, cols); }

int nd4b_dev_qr_f64(int device, void* stream, const double* A, double* Q, double* R,
                    int64_t batch, int rows, int cols, double* workspace, size_t workspace_bytes) {
  if (!A || !Q || !R || batch < 1 || rows < 1 || cols < 1) return fail(ND4B_E_ARG, "dev_qr: bad argument");
  Context* ctx; int sms;
  if (int rc = dev_enter(device, &ctx, &sms)) return rc;
  return check_cuda_launch(nd4b::launch_qr((cudaStream_t)stream, A, Q, R, batch, rows, cols, workspace, workspace_bytes), ctx);
}

int nd4b_dev_tri_solve_f64(int device, void* stream, int op, const double* T, int64_t t_stride, const double* Y, int64_t y_stride,
                           double* X, int64_t batch, int M, int J) {
  if (op < 0 || op > 2 || !T || !Y || !X || batch < 1 || M < 1 || J < 1 || t_stride < 0 || y_stride < 0)
    return fail(ND4B_E_ARG, "dev_tri_solve: bad argument");
  Context* ctx; int sms;
  if (int rc = dev_enter(device, &ctx, &sms)) return rc;
  BatchMap map;
  memset(&map, 0, sizeof map);
  map.a_lin = t_stride;
  map.b_lin = y_stride;
  return check_cuda_launch(nd4b::launch_tri_solve((cudaStream_t)stream, op, T, Y, X, batch, M, J, map), ctx);
}

int nd4b_dev_qr_lstsq_f64(int device, void* stream, const double* Q, const double* R, const double* Y, double* X,
                          int64_t batch, int N, int M, int I, int J) {
  if (!Q || !R || !Y || !X || batch < 1 || N < 1 || M < 1 || I < 1 || J < 1 || M > 32 || I > 32 || I > N)
    return fail(ND4B_E_ARG, "dev_qr_lstsq: bad argument");
  Context* ctx; int sms;
  if (int rc = dev_enter(device, &ctx, &sms)) return rc;
  return check_cuda_launch(nd4b::launch_qr_lstsq((cudaStream_t)stream, Q, R, Y, X, batch, N, M, I, J), ctx);
}

int nd4b_dev_qr_inplace_f64(int device, void* stream, const double* A, const double* Y, double* R, double* QtY,
                            int64_t batch, int M, int N, int L) {
  if (!A || !Y || !R || !QtY || batch < 1 || M < 1 || N < 1 || L < 1) return fail(ND4B_E_ARG, "dev_qr_inplace: bad argument");
  Context* ctx; int sms;
  if (int rc = dev_enter(device, &ctx, &sms)) return rc;
  return check_cuda_launch(nd4b::launch_qr_inplace((cudaStream_t)stream, A, Y, R, QtY, batch, M, N, L), ctx);
}

int nd4b_dev_svd_jac1_f64(int device, void* stream, const double* A, double* U, double* sv, double* V,
                          int64_t batch, int rows, int cols, int* sweeps, double* workspace, size_t workspace_bytes) {
  if (!A || !U || !sv || !V || batch < 1 || rows < 1 || cols < 1) return fail(ND4B_E_ARG, "dev_svd: bad argument");
  Context* ctx; int sms;
  if (int rc = dev_enter(device, &ctx, &sms)) return rc;
  return check_cuda_launch(nd4b::launch_svd_jac1((cudaStream_t)stream, A, U, sv, V, batch, rows, cols, sweeps, nullptr,
                                                 workspace, workspace_bytes), ctx);
}

int nd4b_dev_svd_lstsq_f64(int device, void* stream, const double* U, const double* sv, const double* V, const double* Y, double* X,
                           int64_t batch, int N, int M, int I, int J, int* fail_flag) {
  if (!U || !sv || !V || !Y || !X || batch < 1 || N < 1 || M < 1 || I < 1 || J < 1 || (size_t)M * J * 8 > 200 * 1024)
    return fail(ND4B_E_ARG, "dev_svd_lstsq: bad argument");
  Context* ctx; int sms;
  if (int rc = dev_enter(device, &ctx, &sms)) return rc;
  BatchMap4 map;
  memset(&map, 0, sizeof map);
  map.lin[0] = (int64_t)N * M; map.lin[1] = M; map.lin[2] = (int64_t)M * I; map.lin[3] = (int64_t)N * J;
  return check_cuda_launch(nd4b::launch_svd_lstsq((cudaStream_t)stream, U, sv, V, Y, X, batch, N, M, I, J, map, fail_flag), ctx);
}

int nd4b_dev_svd_sweep_counter(int device, unsigned long long* counter) {
  Context* ctx; int sms;
  if (int rc = dev_enter(device, &ctx, &sms)) return rc;
  nd4b::set_svd_sweep_counter(device, counter);
  return ND4B_OK;
}

int nd4b_dev_svd_pre_sweep_counter(int device, unsigned long long* counter) {
  Context* ctx; int sms;
  if (int rc = dev_enter(device, &ctx, &sms)) return rc;
  nd4b::set_svd_pre_sweep_counter(device, counter);
  return ND4B_OK;
}

// ---- NCCL gather of device-resident shards ("NCCL over NVLink used only to gather results") ----------------------

#define NC(call)                                                                                          \
  do {                                                                                                    \
    ncclResult_t r__ = (call);                                                                            \
    if (r__ != ncclSuccess) return fail(ND4B_E_CUDA, "NCCL error at %s:%d: %s", __FILE__, __LINE__, g_nccl.GetErrorString(r__)); \
  } while (0)

int nd4b_dev_all_gather_f64(const double* const* shards, const int64_t* counts, double* const* full, void* const* streams) {
  if (!shards || !counts || !full) return fail(ND4B_E_ARG, "dev_all_gather: null pointer");
  Context* ctx;
  if (int rc = get_ctx(&ctx)) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  const int nd = (int)ctx->devs.size();
  for (int d = 0; d < nd; d++)
    if (!shards[d] || !full[d] || counts[d] < 0) return fail(ND4B_E_ARG, "dev_all_gather: bad argument for device %d", d);
  if (nd == 1) {   // one device: the shard is the whole array
    CU(cudaSetDevice(ctx->devs[0].id));
    cudaStream_t st = streams ? (cudaStream_t)streams[0] : ctx->devs[0].slots[0].stream;
    if (full[0] != shards[0]) CU(cudaMemcpyAsync(full[0], shards[0], (size_t)counts[0] * 8, cudaMemcpyDeviceToDevice, st));
    return ND4B_OK;
  }
  if (int rc = load_nccl()) return rc;
  if (ctx->comms.empty()) {
    std::vector<int> ids;
    for (auto& dv : ctx->devs) ids.push_back(dv.id);
    ctx->comms.resize(nd);
    ncclResult_t r = g_nccl.CommInitAll(ctx->comms.data(), nd, ids.data());
    if (r != ncclSuccess) { ctx->comms.clear(); return fail(ND4B_E_CUDA, "ncclCommInitAll failed: %s", g_nccl.GetErrorString(r)); }
  }
  // shard r (counts[r] elements on device r) goes to offset sum(counts[0..r)) of every device's full array: one broadcast per
  // shard inside one group — unequal shards (batch not divisible by the device count) need no padding
  NC(g_nccl.GroupStart());
  int64_t off = 0;
  for (int r = 0; r < nd; r++) {
    for (int d = 0; d < nd; d++) {
      cudaStream_t st = streams ? (cudaStream_t)streams[d] : ctx->devs[d].slots[0].stream;
      ncclResult_t res = g_nccl.Broadcast(d == r ? (const void*)shards[r] : (const void*)(full[d] + off), full[d] + off, (size_t)counts[r],
                                          ncclDouble, r, ctx->comms[d], st);
      if (res != ncclSuccess) { g_nccl.GroupEnd(); return fail(ND4B_E_CUDA, "ncclBroadcast failed: %s", g_nccl.GetErrorString(res)); }
    }
    off += counts[r];
  }
  NC(g_nccl.GroupEnd());
  return ND4B_OK;
}

// FP64 pipe probes (not part of the nd.la surface; used by tools/fp64_peak.py and bench.py).
int nd4b_probe_fp64(int device, int which, int iters, int blocks, int threads, float* ms_out) {
  Context* ctx; int sms;
  if (int rc = dev_enter(device, &ctx, &sms)) return rc;
  double* out;
  CU(cudaMalloc(&out, 64));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  auto run = [&]() { return which == 0 ? nd4b::launch_probe_dfma(0, out, iters, blocks, threads) : nd4b::launch_probe_dmma(0, out, iters, blocks, threads); };
  CU(run());
  CU(cudaDeviceSynchronize());
  CU(cudaEventRecord(e0, 0));
  CU(run());
  CU(cudaEventRecord(e1, 0));
  CU(cudaEventSynchronize(e1));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  if (ms_out) *ms_out = ms;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return ND4B_OK;
}

int nd4b_selfcheck_ieee(int device, long long samples, unsigned long long seed, unsigned long long counts[4]) {
  Context* ctx; int sms;
  if (int rc = dev_enter(device, &ctx, &sms)) return rc;
  if (samples <= 0 || !counts) return fail(ND4B_E_ARG, "nd4b_selfcheck_ieee: samples must be positive and counts non-null.");
  unsigned long long* out;
  CU(cudaMalloc(&out, 4 * sizeof(unsigned long long)));
  CU(cudaMemset(out, 0, 4 * sizeof(unsigned long long)));
  const int threads = 256, blocks = sms * 8;
  const long long per_thread = (samples + (long long)threads * blocks - 1) / ((long long)threads * blocks);
  CU(nd4b::launch_selfcheck(0, per_thread, seed, out, blocks, threads));
  CU(cudaMemcpy(counts, out, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  cudaFree(out);
  return ND4B_OK;
}

}  // extern "C"
