// Exchanges between the data-parallel ranks of one node through peer-mapped device memory (NVLink 5 / NVSwitch),
// as plain kernels -- graph-capturable, no NCCL on the step's critical path:
//
//   * allreduce_sgd: every rank reads all ranks' flat f32 prompt gradients (120 KB at K = 24), sums them in rank
//     order (bit-identical on every replica) and applies the SGD(momentum, weight decay) update of both prompt
//     tensors in the same kernel -- the compute step fused with its collective.
//   * gather: a rank pushes its rows of the text features into every peer's buffer (class-sharded text tower,
//     SURVEY.md 8f2; trainers/rpo.py:180-192 computes all of them on every GPU).
//   * reduce_scatter: a rank pulls the peers' gradients of ITS text-feature rows and sums them in f32.
//
// Synchronisation: block b of a kernel meets block b of the same kernel on every peer at the start (everybody's
// inputs are complete, nobody still reads what is about to be overwritten) and at the end (everybody is done with
// the peers' buffers).  Flags are monotonic 32-bit counters in each rank's signal words: flag[channel][src][block],
// written by `src` with st.release.sys, polled locally with ld.acquire.sys.  A peer that does not arrive within
// 60 s traps the kernel (the process fails instead of hanging the GPU).
#include "common.cuh"

namespace rpo {
namespace peer {

constexpr int MAX_WORLD = RPO_PEER_MAX_WORLD;
constexpr int MAX_BLOCKS = RPO_PEER_MAX_BLOCKS;
constexpr int THREADS = 256;

struct Comm {
  uint32_t *sig[MAX_WORLD];  // every rank's signal words as mapped into this process
  uint32_t *epoch;           // this rank's invocation counters [channels][MAX_BLOCKS]
  int rank, world;
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_peer16(const void *p) {
  uint4 v;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer16(void *p, uint4 v) {
  asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ int flag_index(int channel, int src, int block) {
  return (channel * MAX_WORLD + src) * MAX_BLOCKS + block;
}

// every thread of the block calls it; `value` is the same on every rank for this meeting
__device__ __forceinline__ void meet(const Comm &c, int channel, uint32_t value) {
  __threadfence_system();  // this thread's writes (to peers or to local buffers peers will read) before the flag
  __syncthreads();
  if ((int)threadIdx.x < c.world) {
    const int p = threadIdx.x;
    st_release_sys(c.sig[p] + flag_index(channel, c.rank, blockIdx.x), value);
    const uint32_t *mine = c.sig[c.rank] + flag_index(channel, p, blockIdx.x);
    const unsigned long long t0 = globaltimer_ns();
    while ((int32_t)(ld_acquire_sys(mine) - value) < 0) {
      if (globaltimer_ns() - t0 > 60000000000ull) __trap();
    }
  }
  __syncthreads();
}

// invocation number of this (channel, block): counts up once per launch, identically on every rank
__device__ __forceinline__ uint32_t begin_epoch(const Comm &c, int channel) {
  __shared__ uint32_t e;
  if (threadIdx.x == 0) e = c.epoch[channel * MAX_BLOCKS + blockIdx.x];
  __syncthreads();
  return e;
}
__device__ __forceinline__ void end_epoch(const Comm &c, int channel, uint32_t e) {
  if (threadIdx.x == 0) c.epoch[channel * MAX_BLOCKS + blockIdx.x] = e + 1;
}

struct Ptrs {
  void *p[MAX_WORLD];
};

// ---- all-reduce of the flat gradient fused with the SGD update ---------------------------------------------------
// p <- p - lr * buf,  buf <- mom * buf + (gscale * sum_r g_r + wd * p)      (same arithmetic as sgd_kernel)
template <typename T>
__global__ void __launch_bounds__(THREADS) allreduce_sgd_kernel(Comm c, Ptrs grads, T *text_prompt, T *img_prompt,
                                                                long long n_text, long long n, float *buf,
                                                                const float *lr, float mom, float wd, float gscale,
                                                                const int *first) {
  const uint32_t e = begin_epoch(c, 0);
  meet(c, 0, 2 * e + 1);
  const float step = *lr;
  const bool is_first = first && *first;
  for (long long i = ((long long)blockIdx.x * THREADS + threadIdx.x) * 4; i < n; i += (long long)gridDim.x * THREADS * 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < c.world; ++r) {
      uint4 v = ld_peer16((const float *)grads.p[r] + i);
      acc[0] += __uint_as_float(v.x);
      acc[1] += __uint_as_float(v.y);
      acc[2] += __uint_as_float(v.z);
      acc[3] += __uint_as_float(v.w);
    }
    T *p = i < n_text ? text_prompt + i : img_prompt + (i - n_text);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float pv = tof<T>(p[j]);
      float gv = gscale * acc[j] + wd * pv;
      float b = gv;
      if (mom != 0.f) {
        b = is_first ? gv : mom * buf[i + j] + gv;
        buf[i + j] = b;
      }
      p[j] = fromf<T>(pv - step * b);
    }
  }
  meet(c, 0, 2 * e + 2);  // nobody rewrites its gradient buffer while a peer still reads it
  end_epoch(c, 0, e);
}

// ---- all-gather by pushing: bytes [off, off + nbytes) of this rank's buffer into every peer's buffer -------------
__global__ void __launch_bounds__(THREADS) gather_kernel(Comm c, Ptrs bufs, long long off, long long nbytes) {
  const uint32_t e = begin_epoch(c, 1);
  meet(c, 1, 2 * e + 1);  // every peer is past the last reader of its buffer
  const char *src = (const char *)bufs.p[c.rank] + off;
  for (long long i = ((long long)blockIdx.x * THREADS + threadIdx.x) * 16; i < nbytes; i += (long long)gridDim.x * THREADS * 16) {
    const uint4 v = *reinterpret_cast<const uint4 *>(src + i);
    for (int r = 0; r < c.world; ++r)
      if (r != c.rank) st_peer16((char *)bufs.p[r] + off + i, v);
  }
  meet(c, 1, 2 * e + 2);  // all rows of all ranks have landed here
  end_epoch(c, 1, e);
}

// ---- reduce-scatter by pulling: elements [off, off + n) of this rank's buffer <- sum over ranks, f32 accumulation -
template <typename T>
__global__ void __launch_bounds__(THREADS) reduce_scatter_kernel(Comm c, Ptrs bufs, long long off, long long n) {
  constexpr int V = 16 / sizeof(T);
  const uint32_t e = begin_epoch(c, 2);
  meet(c, 2, 2 * e + 1);  // every rank's buffer is complete
  for (long long i = ((long long)blockIdx.x * THREADS + threadIdx.x) * V; i < n; i += (long long)gridDim.x * THREADS * V) {
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.f;
    for (int r = 0; r < c.world; ++r) {
      uint4 raw = ld_peer16((const T *)bufs.p[r] + off + i);
      const T *v = reinterpret_cast<const T *>(&raw);
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] += tof<T>(v[j]);
    }
    uint4 out;
    T *o = reinterpret_cast<T *>(&out);
#pragma unroll
    for (int j = 0; j < V; ++j) o[j] = fromf<T>(acc[j]);
    *reinterpret_cast<uint4 *>((T *)bufs.p[c.rank] + off + i) = out;
  }
  meet(c, 2, 2 * e + 2);  // peers are done reading this rank's buffer
  end_epoch(c, 2, e);
}

static int make_comm(const RpoPeerComm *pc, Comm &c) {
  RPO_REQUIRE(pc && pc->epoch, "null argument");
  RPO_REQUIRE(pc->world >= 2 && pc->world <= MAX_WORLD && pc->rank >= 0 && pc->rank < pc->world, "rank / world");
  for (int r = 0; r < pc->world; ++r) {
    RPO_REQUIRE(pc->signals[r], "null signal pointer");
    c.sig[r] = (uint32_t *)pc->signals[r];
  }
  c.epoch = (uint32_t *)pc->epoch;
  c.rank = pc->rank;
  c.world = pc->world;
  return RPO_OK;
}

static int make_ptrs(void *const *bufs, int world, Ptrs &p) {
  RPO_REQUIRE(bufs, "null argument");
  for (int r = 0; r < world; ++r) {
    RPO_REQUIRE(bufs[r] && ((uintptr_t)bufs[r] & 15) == 0, "peer buffers must be non-null and 16-byte aligned");
    p.p[r] = bufs[r];
  }
  return RPO_OK;
}

static int blocks_for(long long chunks) {
  long long b = (chunks + THREADS - 1) / THREADS;
  return (int)(b < 1 ? 1 : (b > MAX_BLOCKS ? MAX_BLOCKS : b));
}

}  // namespace peer
}  // namespace rpo

using namespace rpo;
using namespace rpo::peer;

extern "C" {

size_t rpo_peer_signal_bytes(void) { return sizeof(uint32_t) * 3 * MAX_WORLD * MAX_BLOCKS; }
size_t rpo_peer_epoch_bytes(void) { return sizeof(uint32_t) * 3 * MAX_BLOCKS; }

int rpo_peer_allreduce_sgd(const RpoPeerComm *comm, void *const *grad_flat, void *text_prompt, void *img_prompt,
                           int32_t dtype, int64_t n_text, int64_t n_total, float *momentum_buf, const float *lr,
                           float momentum, float weight_decay, float grad_scale, const int32_t *first_step,
                           void *stream) {
  Comm c{};
  Ptrs g{};
  RPO_TRY(make_comm(comm, c));
  RPO_TRY(make_ptrs(grad_flat, c.world, g));
  RPO_REQUIRE(text_prompt && img_prompt && lr, "null argument");
  RPO_REQUIRE(momentum == 0.f || momentum_buf, "momentum needs a buffer");
  RPO_REQUIRE(n_text > 0 && n_total > n_text && n_text % 4 == 0 && n_total % 4 == 0, "element counts must be multiples of 4");
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = blocks_for(n_total / 4);
#define LAUNCH(T)                                                                                                     \
  allreduce_sgd_kernel<T><<<nb, THREADS, 0, st>>>(c, g, (T *)text_prompt, (T *)img_prompt, n_text, n_total,            \
                                                  momentum_buf, lr, momentum, weight_decay, grad_scale, first_step)
  switch (dtype) {
    case RPO_F32: LAUNCH(float); break;
    case RPO_F16: LAUNCH(__half); break;
    case RPO_BF16: LAUNCH(__nv_bfloat16); break;
    default: set_error("invalid dtype"); return RPO_ERR_INVALID;
  }
#undef LAUNCH
  RPO_LAUNCH_CHECK();
  return RPO_OK;
}

int rpo_peer_all_gather(const RpoPeerComm *comm, void *const *bufs, int64_t offset_bytes, int64_t nbytes,
                        int64_t nbytes_max, void *stream) {
  Comm c{};
  Ptrs b{};
  RPO_TRY(make_comm(comm, c));
  RPO_TRY(make_ptrs(bufs, c.world, b));
  RPO_REQUIRE(offset_bytes >= 0 && nbytes >= 0 && nbytes <= nbytes_max && offset_bytes % 16 == 0 && nbytes % 16 == 0,
              "byte ranges must be multiples of 16");
  cudaStream_t st = (cudaStream_t)stream;
  // the grid follows the LARGEST part so that every rank launches the same number of blocks
  gather_kernel<<<blocks_for(nbytes_max / 16 / 4 + 1), THREADS, 0, st>>>(c, b, offset_bytes, nbytes);
  RPO_LAUNCH_CHECK();
  return RPO_OK;
}

int rpo_peer_reduce_scatter(const RpoPeerComm *comm, void *const *bufs, int32_t dtype, int64_t offset_elems,
                            int64_t n_elems, int64_t n_elems_max, void *stream) {
  Comm c{};
  Ptrs b{};
  RPO_TRY(make_comm(comm, c));
  RPO_TRY(make_ptrs(bufs, c.world, b));
  const int V = 16 / (int)dtype_size(dtype);
  RPO_REQUIRE(offset_elems >= 0 && n_elems >= 0 && n_elems <= n_elems_max && offset_elems % V == 0 && n_elems % V == 0,
              "element ranges must be multiples of 16 bytes");
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = blocks_for(n_elems_max / V / 2 + 1);
  switch (dtype) {
    case RPO_F32: reduce_scatter_kernel<float><<<nb, THREADS, 0, st>>>(c, b, offset_elems, n_elems); break;
    case RPO_F16: reduce_scatter_kernel<__half><<<nb, THREADS, 0, st>>>(c, b, offset_elems, n_elems); break;
    case RPO_BF16: reduce_scatter_kernel<__nv_bfloat16><<<nb, THREADS, 0, st>>>(c, b, offset_elems, n_elems); break;
    default: set_error("invalid dtype"); return RPO_ERR_INVALID;
  }
  RPO_LAUNCH_CHECK();
  return RPO_OK;
}

}  // extern "C"
