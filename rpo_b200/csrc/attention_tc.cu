// tcgen05 forms of the read-only masked attention forward for the vision tower (every group has the same
// number n of context rows, no causal mask): clip/model.py:186 under visual_mask of trainers/rpo.py:153-159.
// The metric's masked-attention kernel.
//
// One CTA per SM walks a contiguous range of (image, head, 128-query tile) work items.  The query rows of an (image,
// head) are its n context rows followed by its K prompt rows (the last tile holds both); keys / values are the n
// context rows only -- that IS the mask.  K and V of a head are loaded once and serve all its query tiles.
//
// Two kernels share the producer and the work decomposition:
//   * ro_attn_fwd_pp (second half of the file): n <= 224 keys (ViT-B/16: 197).  Two S slots in tensor memory, two
//     softmax warp groups on alternate tiles, probabilities kept in tensor memory.  Described at its definition.
//   * ro_attn_fwd_tc (below): 225 .. 272 keys (ViT-L/14: 257), ONE 512-column slot, S in two UMMA N blocks (144 + rest).
//
// ro_attn_fwd_tc:
//   warp 0 (one thread)   TMA producer: K | V of the next (image, head) through a ring of three buffers, Q tiles into a
//                         ring (cp.async.bulk.tensor, 128B swizzle, straight out of the [rows, 3D] q|k|v matrix and
//                         the [G*K, D] prompt-q matrix).
//   warp 1 (one thread)   S = Q K^T (M=128, two N blocks per 16-wide k step, f32 in tensor memory)
//   warp 2 (one thread)   O = P V, one tile behind: P from a 128B-swizzled shared-memory tile, V consumed as loaded
//                         (MN-major B operand), into the last 64 TMEM columns
//   warps 4..19           softmax, four threads per query row (TMEM lane), each with a quarter of the 16-key blocks held
//                         in registers: row maximum (quarters exchanged through shared memory),
//                         p = ex2((s - max) / 8 log2 e), f32 row sum, p rounded to the dtype into the P tile; a block's
//                         registers take the next tile's scores as soon as its probabilities are stored
//   warps 20..23          epilogue, one thread per query row: O out of TMEM, times 1/sum, 128 contiguous bytes to global
// Registers: the roles re-partition the register file with setmaxnreg (softmax threads hold up to 5 x 16 scores).
#include <stdlib.h>

#include <map>
#include <type_traits>

#include "common.cuh"
#include "tc_common.cuh"

namespace rpo {

namespace atc {

using namespace tc;

static constexpr int HD = 64;
static constexpr int ROW_BYTES = HD * 2;
static constexpr int QT = 128;                 // query rows per tile = UMMA M
static constexpr int Q_TILE_BYTES = QT * ROW_BYTES;
static constexpr int PROD_WARPS = 4, SM_WARPS = 16, EPI_WARPS = 4;
static constexpr int PARTS = SM_WARPS / 4;  // softmax threads per query row: each takes a quarter of the key blocks
static constexpr int THREADS = 32 * (PROD_WARPS + SM_WARPS + EPI_WARPS);
static constexpr int TMEM_COLS = 512;
// The CTA is launched with 80 registers per thread (768 threads); setmaxnreg moves registers between the roles. What
// the softmax warps gain must come out of the CTA's own pool, i.e. out of what the other roles give back:
// 512 x (104 - 80) = 128 x (80 - 24) + 128 x (80 - 40).
// 512 x (S - 80) = 128 x (80 - P) + 128 x (80 - E):  4 blocks per softmax thread (ViT-B/16): S 96, P 56, E 40;
// 5 blocks (ViT-L/14): S 104, P 24, E 40.
static constexpr int REGS_EPI = 40;
template <int MAXB>
struct Regs {
  static constexpr int SOFTMAX = MAXB <= 4 ? 96 : 104;
  static constexpr int PROD = MAXB <= 4 ? 56 : 24;
};
static constexpr int MAX_Q_RING = 4;
static constexpr int MAX_ROUNDS = 5;            // key blocks per softmax thread, at most
static constexpr int KV_RING = 3;               // K and V tiles rotate through three buffers (see the kernel)
static constexpr int P_TILE_BYTES = QT * ROW_BYTES;  // P as A operand: 128 x 64-key tiles, K-major, 128B swizzle (like Q)

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <typename T>
__device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <>
__device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&v);
}
template <>
__device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&v);
}

// explicit shared-space accesses with 32-bit addresses (a generic pointer into the dynamic shared memory makes the
// compiler emit generic ST.E / LD.E and carry 64-bit addresses through the softmax loop)
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
// named barrier of the PARTS warps that share a TMEM lane quarter
__device__ __forceinline__ void quad_bar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(PARTS * 32) : "memory"); }
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

#ifdef RPO_DIAG
// phase timestamps of CTA 0 (tools/attn_trace.py): trace[tile * 16 + event] = SM clock
__device__ long long *g_attn_trace = nullptr;
#define ATTN_TRACE(j, ev)                                                                  \
  do {                                                                                     \
    if (blockIdx.x == 0 && trace_buf && (j) < 64) trace_buf[(j) * 16 + (ev)] = clock64(); \
  } while (0)
// wall-clock stamps of EVERY CTA behind the phase table: trace[1024 + 4 * cta + k] = globaltimer (ns) at
// k = 0 kernel entry, 1 dependency released (producer past griddepcontrol.wait), 2 all roles done, 3 exit
#define ATTN_WALL(k)                                                                                   \
  do {                                                                                                 \
    if (trace_buf) {                                                                                   \
      unsigned long long t_;                                                                           \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                           \
      trace_buf[1024 + 4 * blockIdx.x + (k)] = (long long)t_;                                          \
    }                                                                                                  \
  } while (0)
#else
#define ATTN_TRACE(j, ev) \
  do {                    \
  } while (0)
#define ATTN_WALL(k) \
  do {               \
  } while (0)
#endif

struct Geo {
  int n;            // context rows (keys) per group
  int n16;          // keys rounded up to the UMMA K step
  int nblk;         // n16 / 16
  int K;            // prompt rows per group
  int prompt_tile;  // query tile that holds the prompt rows ...
  int prompt_row;   // ... starting at this row of the tile (== context rows in that tile)
  int H;
  int tiles;        // query tiles per (group, head)
  int q_ring;       // Q tile buffers
  int n_first;      // keys of the first N block of S (the second holds n16 - n_first; 0 columns if n16 <= 256)
};

struct Item {
  int unit, t, h, g, c_rows, p_rows;
};
__device__ __forceinline__ Item item_of(const Geo &geo, int id) {
  Item it;
  it.unit = id / geo.tiles;
  it.t = id - it.unit * geo.tiles;
  it.h = it.unit % geo.H;
  it.g = it.unit / geo.H;
  it.c_rows = min(max(geo.n - it.t * QT, 0), QT);     // context query rows of this tile
  it.p_rows = (it.t == geo.prompt_tile) ? geo.K : 0;   // prompt query rows, right behind them
  return it;
}

// the next work item without divisions (the loops of the softmax / epilogue warps run once per tile)
__device__ __forceinline__ void advance(const Geo &geo, Item &it) {
  if (++it.t == geo.tiles) {
    it.t = 0;
    ++it.unit;
    if (++it.h == geo.H) {
      it.h = 0;
      ++it.g;
    }
  }
  it.c_rows = min(max(geo.n - it.t * QT, 0), QT);
  it.p_rows = (it.t == geo.prompt_tile) ? geo.K : 0;
}

// MAXB: 16-key blocks a softmax thread keeps in registers (>= ceil(nblk / PARTS))
template <typename T, int MAXB>
__global__ void __launch_bounds__(THREADS, 1)
    ro_attn_fwd_tc(const __grid_constant__ CUtensorMap map_full,   // q|k|v matrix, box 64 x 128
                   const __grid_constant__ CUtensorMap map_kvt,    // q|k|v matrix, box 64 x (n16 % 128)
                   const __grid_constant__ CUtensorMap map_qt,     // q|k|v matrix, box 64 x (n % 128)
                   const __grid_constant__ CUtensorMap map_prompt, // prompt-q matrix, box 64 x K
                   T *__restrict__ out_ctx, T *__restrict__ out_prompt, Geo geo, int total_items) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // tensor memory: S slot s at columns [s * SLOT_COLS, ...), O in the last 64 columns
  constexpr int SLOTS = 1;  // one S slot (the two-slot form is ro_attn_fwd_pp); the slot arithmetic below keeps the general form
  constexpr int SLOT_COLS = SLOTS == 2 ? 224 : 0;
  constexpr uint32_t O_COL = TMEM_COLS - HD;
  const int n = geo.n, n16 = geo.n16, K = geo.K, H = geo.H, NQ = geo.q_ring;
  const int D = H * HD;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv_bytes = n16 * ROW_BYTES;
  // shared memory: [K|V ring of 3 | P | Q ring | barriers | row maxima | row sums]
  const uint32_t kv_base = smem_u32(smem);
  const uint32_t p_base = kv_base + KV_RING * kv_bytes;
  const int p_bytes = ((geo.nblk + 3) / 4) * P_TILE_BYTES;  // 64-key tiles
  const uint32_t q_base = p_base + p_bytes;
  uint8_t *tail = smem + KV_RING * kv_bytes + p_bytes + NQ * Q_TILE_BYTES;
  uint64_t *bars = reinterpret_cast<uint64_t *>(tail);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int idx) { return bar0 + 8u * idx; };
  // K or V landed / its ring entry free again; Q landed / Q slot free; S complete / S in registers (slot free);
  // P stored / P consumed; O complete / O read
  // (P consumed: one barrier per ROUND of the P V issue order, see the P V issuer)
  constexpr int B_KVFULL = 0, B_KVFREE = 3, B_QFULL = 6, B_QFREE = 10, B_SFULL = 14, B_SFREE = 16, B_PFULL = 18,
                B_OFULL = 19, B_OFREE = 20, B_PFREE = 21, N_BARS = 21 + MAX_ROUNDS;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + N_BARS);
  const uint32_t red_max = smem_u32(bars + N_BARS + 1);        // f32 [2][PARTS][QT]  (tile parity, key part, row)
  // (two buffers are enough: the softmax of tile j+2 writes its sums after the P V of tile j+1 has been issued, which
  // waited for the epilogue of tile j to have read O -- and with it these sums)
  const uint32_t red_sum = red_max + 2 * PARTS * QT * 4;        // f32 [2][PARTS][QT]  (tile parity, key part, row)

  // this CTA's contiguous range of work items: consecutive tiles of an (image, head) share its K / V
  const int item0 = (int)((long long)total_items * blockIdx.x / gridDim.x);
  const int item1 = (int)((long long)total_items * (blockIdx.x + 1) / gridDim.x);
  const int count = item1 - item0;
#ifdef RPO_DIAG
  long long *const trace_buf = g_attn_trace;
  if (threadIdx.x == 0) ATTN_TRACE(0, 12);
#endif

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_full)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_kvt)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_qt)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_prompt)) : "memory");
    for (int i = 0; i < KV_RING; ++i) {
      mbar_init(BAR(B_KVFULL + i), 1);
      mbar_init(BAR(B_KVFREE + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(BAR(B_SFULL + i), 1);
      mbar_init(BAR(B_SFREE + i), SM_WARPS);
    }
    mbar_init(BAR(B_PFULL), SM_WARPS);
    for (int i = 0; i < MAX_ROUNDS; ++i) mbar_init(BAR(B_PFREE + i), 1);
    mbar_init(BAR(B_OFULL), 1);
    mbar_init(BAR(B_OFREE), EPI_WARPS);
    for (int i = 0; i < MAX_Q_RING; ++i) {
      mbar_init(BAR(B_QFULL + i), 1);
      mbar_init(BAR(B_QFREE + i), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  if (threadIdx.x == 0) ATTN_TRACE(0, 13);

  // unit (= image, head) bookkeeping of tile j of this CTA: index of its unit within the range, first / last tile of
  // that unit inside the range.  K of unit u lives in ring entry (2u) % 3, V in (2u + 1) % 3: V of unit u+1 takes over
  // the entry of K of unit u (dead after the unit's last S), K of unit u+2 the entry of V of unit u.
  const int tiles = geo.tiles;
  const int unit0 = item0 / tiles;
  auto unit_idx = [&](int j) { return (item0 + j) / tiles - unit0; };
  auto first_of_unit = [&](int j) { return j == 0 || (item0 + j) % tiles == 0; };
  auto last_of_unit = [&](int j) { return j + 1 == count || (item0 + j + 1) % tiles == 0; };
  // the c-th use (c = 0, 1, ...) of ring entry e = c-th element of {2u, 2u+1 : ...} congruent to e: element index x
  // (x = 2u for K, 2u + 1 for V) is use number x / 3 of entry x % 3
  if (warp < PROD_WARPS) {
    setmaxnreg_dec<Regs<MAXB>::PROD>();
    if (warp == 0 && lane == 0) {
      // ===== TMA producer =====
      pdl_wait();
      for (int j = 0; j < count; ++j) {
        const Item it = item_of(geo, item0 + j);
        if (first_of_unit(j)) {
          const int u = unit_idx(j);
#pragma unroll
          for (int kv = 0; kv < 2; ++kv) {  // K, then V
            const int x = 2 * u + kv, e = x % KV_RING, use = x / KV_RING;
            if (use >= 1) mbar_wait(BAR(B_KVFREE + e), (uint32_t)((use - 1) & 1));
            const uint32_t dst = kv_base + e * kv_bytes;
            mbar_arrive_expect_tx(BAR(B_KVFULL + e), (uint32_t)kv_bytes);
            for (int r = 0; r < n16; r += 128)
              tma_load_2d(dst + r * ROW_BYTES, (n16 - r >= 128) ? &map_full : &map_kvt, BAR(B_KVFULL + e),
                          (1 + kv) * D + it.h * HD, it.g * n + r);
          }
          ATTN_TRACE(j, 9);
        }
        const int qs = j % NQ;
        if (j >= NQ) mbar_wait(BAR(B_QFREE + qs), (uint32_t)(((j / NQ) - 1) & 1));
        const uint32_t Qs = q_base + qs * Q_TILE_BYTES;
        mbar_arrive_expect_tx(BAR(B_QFULL + qs), (uint32_t)((it.c_rows + it.p_rows) * ROW_BYTES));
        if (it.c_rows == QT)
          tma_load_2d(Qs, &map_full, BAR(B_QFULL + qs), it.h * HD, it.g * n + it.t * QT);
        else if (it.c_rows > 0)
          tma_load_2d(Qs, &map_qt, BAR(B_QFULL + qs), it.h * HD, it.g * n + it.t * QT);
        if (it.p_rows > 0)
          tma_load_2d(Qs + geo.prompt_row * ROW_BYTES, &map_prompt, BAR(B_QFULL + qs), it.h * HD, it.g * K);
        ATTN_TRACE(j, 10);
      }
    } else if (warp == 1 && elect_one()) {
      // ===== S = Q K^T issuer =====
      const uint32_t fmt = Num<T>::dtype == RPO_BF16 ? 1u : 0u;
      const uint32_t idesc_s0 = make_idesc((int)fmt, QT, geo.n_first);
      const uint32_t idesc_s1 = make_idesc((int)fmt, QT, n16 - geo.n_first > 0 ? n16 - geo.n_first : 16);
      const bool two_blocks = n16 > geo.n_first;
      for (int j = 0; j < count; ++j) {
        const int slot = j % SLOTS, qs = j % NQ;
        const int x = 2 * unit_idx(j), e = x % KV_RING;
        mbar_wait(BAR(B_QFULL + qs), (uint32_t)((j / NQ) & 1));
        if (first_of_unit(j)) mbar_wait(BAR(B_KVFULL + e), (uint32_t)((x / KV_RING) & 1));
        ATTN_TRACE(j, 11);
        // the slot's previous S is in the softmax threads' registers
        if (j >= SLOTS) mbar_wait(BAR(B_SFREE + slot), (uint32_t)(((j / SLOTS) - 1) & 1));
        tc_fence_after();
        const uint32_t Qs = q_base + qs * Q_TILE_BYTES, Ks = kv_base + e * kv_bytes;
        const uint32_t tslot = tmem_base + (uint32_t)(slot * SLOT_COLS);
        const uint64_t adesc = make_smem_desc(Qs), bdesc0 = make_smem_desc(Ks);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_f16(tslot, adesc + 2u * k, bdesc0 + 2u * k, idesc_s0, k != 0);
        if (two_blocks) {
          const uint64_t bdesc1 = make_smem_desc(Ks + geo.n_first * ROW_BYTES);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            umma_f16(tslot + (uint32_t)geo.n_first, adesc + 2u * k, bdesc1 + 2u * k, idesc_s1, k != 0);
        }
        umma_commit(BAR(B_SFULL + slot));
        umma_commit(BAR(B_QFREE + qs));
        if (last_of_unit(j)) umma_commit(BAR(B_KVFREE + e));  // K of the unit is dead once this S is complete
        ATTN_TRACE(j, 0);
      }
    } else if (warp == 2 && elect_one()) {
      // ===== O = P V issuer: A = P from shared memory (K-major 128 x 16 blocks, 32B swizzle), B = V as loaded =====
      const uint32_t fmt = Num<T>::dtype == RPO_BF16 ? 1u : 0u;
      const uint32_t idesc_o = make_idesc((int)fmt, QT, HD) | (1u << 16);  // B (= V) is MN-major
      const uint64_t pdesc = make_smem_desc(p_base);
      // Issue order: round r takes the r-th key block of each of the PARTS softmax threads of a row.  The softmax
      // threads write their blocks in the same order, so a thread may overwrite its r-th block of the single P buffer
      // as soon as round r of the previous tile has been consumed (B_PFREE + r) instead of waiting for the whole P V.
      // Uniform arithmetic on kernel parameters only (no table, no vector-register operands).
      const int base_nb = geo.nblk / PARTS, extra = geo.nblk % PARTS;
      const int rounds = base_nb + (extra ? 1 : 0);
      for (int j = 0; j < count; ++j) {
        const int x = 2 * unit_idx(j) + 1, e = x % KV_RING;
        mbar_wait(BAR(B_PFULL), (uint32_t)(j & 1));
        ATTN_TRACE(j, 1);
        if (first_of_unit(j)) mbar_wait(BAR(B_KVFULL + e), (uint32_t)((x / KV_RING) & 1));
        if (j >= 1) mbar_wait(BAR(B_OFREE), (uint32_t)((j - 1) & 1));  // the previous tile's O is out of tensor memory
        tc_fence_after();
        const uint64_t vdesc = make_smem_desc(kv_base + e * kv_bytes);
        uint32_t acc = 0;
        for (int r = 0; r < rounds; ++r) {
#pragma unroll
          for (int pt = 0; pt < PARTS; ++pt) {
            const int nbp = base_nb + (pt < extra ? 1 : 0);
            if (r < nbp) {
              // key block blk: 32 bytes along K inside the 128-byte row of P tile blk / 4, 2 KB of V
              const int blk = pt * base_nb + (pt < extra ? pt : extra) + r;
              umma_f16(tmem_base + O_COL, pdesc + (uint64_t)((blk >> 2) * (P_TILE_BYTES >> 4) + 2 * (blk & 3)),
                       vdesc + (uint64_t)(blk * 128), idesc_o, acc);
              acc = 1;
            }
          }
          umma_commit(BAR(B_PFREE + r));
        }
        umma_commit(BAR(B_OFULL));
        if (last_of_unit(j)) umma_commit(BAR(B_KVFREE + e));  // V of the unit is dead once this O is complete
        ATTN_TRACE(j, 2);
      }
    }
  } else if (warp < PROD_WARPS + SM_WARPS) {
    // ===== softmax: warps w, w+4, w+8, w+12 own TMEM lanes (= query rows) 32*(w%4) .. +31 and split the key blocks =====
    // Four warps per scheduler hide each other's MUFU / TMEM latencies.  Software-pipelined over tiles: a block's
    // registers take the next tile's scores as soon as its probabilities are stored, so the tensor-memory reads (the
    // scarcest resource: 64 B/clk per SM) run beside the exponentials.
    setmaxnreg_inc<Regs<MAXB>::SOFTMAX>();
    const int q = warp & 3;
    const int part = (warp - PROD_WARPS) >> 2;
    const int row = q * 32 + lane;
    const bool tracer = threadIdx.x == PROD_WARPS * 32;
    const float sl2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    const int nblk = geo.nblk;
    // this thread's 16-key blocks [b0, b0 + nb): the first (nblk % PARTS) parts hold one block more.  The kernel is
    // instantiated with MAXB = the larger count, so blocks 0 .. MAXB-2 exist for every thread and only the last one is
    // conditional (has_last): the hot loop carries one warp-uniform branch instead of one per block.
    const int base_nb = nblk / PARTS, extra = nblk % PARTS;
    const int b0 = part * base_nb + min(part, extra);
    const bool has_last = (base_nb + (part < extra ? 1 : 0)) == MAXB;
    // columns >= n are padding: they sit at the end of the last key block, which the last part holds
    const bool pads = part == PARTS - 1 && n < n16;
    const int pad_first = n - (nblk - 1) * 16;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b0 * 16);
    // P as the A operand of P V: 128 x 64-key tiles, K-major, 128-byte rows, 16-byte chunk index ^ (row & 7) -- the
    // layout TMA gives the Q tile.  Key block g of this row = chunks 2 (g & 3), 2 (g & 3) + 1 of row `row` of tile g / 4.
    const uint32_t p_row = p_base + (uint32_t)(row * ROW_BYTES);
    const uint32_t sw = (uint32_t)(row & 7);
    // one S slot: the next S is issued when this one is in registers, poll for it half way through the blocks
    constexpr int PF_AT = SLOTS == 2 ? 0 : MAXB / 2;
    uint32_t s[MAXB][16];
    auto load_scores = [&](uint32_t taddr) {  // all of this thread's blocks; completion: wait_scores()
#pragma unroll
      for (int b = 0; b < MAXB - 1; ++b) tmem_ld16_nowait(taddr + (uint32_t)(b * 16), s[b]);
      if (has_last) tmem_ld16_nowait(taddr + (uint32_t)((MAXB - 1) * 16), s[MAXB - 1]);
    };
    auto wait_scores = [&]() {
#pragma unroll
      for (int b = 0; b < MAXB; ++b) tmem_ld_wait(s[b]);
    };
    auto mask_pad = [&](uint32_t (&blk)[16]) {
#pragma unroll
      for (int e = 0; e < 16; ++e)
        if (e >= pad_first) blk[e] = 0xff800000u;  // -inf: drops out of the maximum, ex2 gives 0
    };
    bool loaded = false;  // s[] holds the scores of the tile about to be processed
    Item it = item_of(geo, item0), nx = it;
    for (int j = 0; j < count; ++j, it = nx) {
      advance(geo, nx);
      const int slot = j % SLOTS;
      const bool valid = q * 32 < it.c_rows + it.p_rows;  // warp-uniform, identical for the warps of a lane quarter
      if (!loaded) {
        // every warp waits for S and arrives on "S in registers", also warps without valid rows
        mbar_wait(BAR(B_SFULL + slot), (uint32_t)((j / SLOTS) & 1));
        if (valid) {
          tc_fence_after();
          load_scores(lane_base + (uint32_t)(slot * SLOT_COLS));
          wait_scores();
          tc_fence_before();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_SFREE + slot));
      }
      loaded = false;
      if (tracer) ATTN_TRACE(j, 4);
      const bool have_next = j + 1 < count;
      const bool nvalid = have_next && q * 32 < nx.c_rows + nx.p_rows;
      const int nslot = (j + 1) % SLOTS;
      const uint32_t npar = (uint32_t)(((j + 1) / SLOTS) & 1);
      const uint32_t ntaddr = lane_base + (uint32_t)(nslot * SLOT_COLS);
      bool pf = false;  // the next tile's scores are being prefetched into the freed registers
      if (valid) {
        if (pads) {
          if (has_last)
            mask_pad(s[MAXB - 1]);
          else if (MAXB >= 2)
            mask_pad(s[MAXB >= 2 ? MAXB - 2 : 0]);
        }
        // ---- row maximum over this thread's keys (four independent chains) ----
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        auto max_block = [&](const uint32_t (&blk)[16]) {
#pragma unroll
          for (int e = 0; e < 16; e += 4)
            m4[e >> 2] = fmaxf(m4[e >> 2], fmaxf(fmaxf(__uint_as_float(blk[e]), __uint_as_float(blk[e + 1])),
                                                 fmaxf(__uint_as_float(blk[e + 2]), __uint_as_float(blk[e + 3]))));
        };
#pragma unroll
        for (int b = 0; b < MAXB - 1; ++b) max_block(s[b]);
        if (has_last) max_block(s[MAXB - 1]);
        float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        const uint32_t rmax = red_max + (uint32_t)(((j & 1) * PARTS * QT + row) * 4);
        sts_f32(rmax + (uint32_t)(part * QT * 4), mx);
        quad_bar_sync(1 + q);
        mx = fmaxf(fmaxf(lds_f32(rmax), lds_f32(rmax + QT * 4)), fmaxf(lds_f32(rmax + 2 * QT * 4), lds_f32(rmax + 3 * QT * 4)));
        if (tracer) ATTN_TRACE(j, 5);
        const float off = mx * sl2;  // finite: every row sees key 0
        // ---- probabilities -> P blocks in shared memory, row sum ----
        float l0 = 0.f, l1 = 0.f;
        auto prob_block = [&](int b, uint32_t (&blk)[16]) {  // b: compile-time constant at every call site
          // the single P buffer: round b of the previous tile's P V has consumed this block
          if (j >= 1) mbar_wait(BAR(B_PFREE + b), (uint32_t)((j - 1) & 1));
          if (tracer && b == 0 && j > 0) ATTN_TRACE(j, 14);
          uint32_t pk[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(blk[2 * e]), sl2, -off));
            const float p1 = ex2_approx(fmaf(__uint_as_float(blk[2 * e + 1]), sl2, -off));
            l0 += p0;
            l1 += p1;
            pk[e] = pack2<T>(p0, p1);
          }
          const uint32_t g = (uint32_t)(b0 + b), tile = p_row + (g >> 2) * P_TILE_BYTES, c = 2u * (g & 3u);
          sts128(tile + ((c ^ sw) << 4), pk[0], pk[1], pk[2], pk[3]);
          sts128(tile + (((c + 1u) ^ sw) << 4), pk[4], pk[5], pk[6], pk[7]);
          if (pf) tmem_ld16_nowait(ntaddr + (uint32_t)(b * 16), blk);
          if (tracer && b == 0 && j > 0) ATTN_TRACE(j, 15);
        };
#pragma unroll
        for (int b = 0; b < MAXB; ++b) {
          if (b == PF_AT && nvalid) {
            if (SLOTS == 2) {
              mbar_wait(BAR(B_SFULL + nslot), npar);  // issued long ago: the slot was free as soon as it was read
              pf = true;
            } else {
              pf = __all_sync(0xffffffffu, mbar_try_wait(BAR(B_SFULL + nslot), npar));
            }
            if (pf) {
              tc_fence_after();
#pragma unroll
              for (int c = 0; c < PF_AT; ++c) tmem_ld16_nowait(ntaddr + (uint32_t)(c * 16), s[c]);
            }
          }
          if (b < MAXB - 1)
            prob_block(b, s[b]);
          else if (has_last)
            prob_block(b, s[b]);
        }
        sts_f32(red_sum + (uint32_t)((((j & 1) * PARTS + part) * QT + row) * 4), l0 + l1);
        if (tracer && j > 0) ATTN_TRACE(j, 9);
        // make the generic-proxy stores of P visible to the tensor core (async proxy)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      } else if (j >= 1) {
        // a warp without valid rows must not arrive for this tile while the previous tile's phase of "P stored" is
        // still open: the previous P V (which waited for that phase) has at least started
        mbar_wait(BAR(B_PFREE), (uint32_t)((j - 1) & 1));
      }
      __syncwarp();
      if (tracer) ATTN_TRACE(j, 6);
      if (lane == 0) mbar_arrive(BAR(B_PFULL));
      if (tracer && j > 0) ATTN_TRACE(j, 12);
      if (nvalid) {
        if (!pf) {
          mbar_wait(BAR(B_SFULL + nslot), npar);
          tc_fence_after();
          load_scores(ntaddr);
        }
        wait_scores();
        if (tracer && j > 0) ATTN_TRACE(j, 13);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_SFREE + nslot));
        loaded = true;
        if (tracer) ATTN_TRACE(j + 1, 3);
      }
    }
  } else {
    // ===== epilogue: one thread per query row =====
    setmaxnreg_dec<REGS_EPI>();
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool tracer = threadIdx.x == (PROD_WARPS + SM_WARPS) * 32;
    const uint32_t taddr = tmem_base + O_COL + ((uint32_t)(q * 32) << 16);
    pdl_wait();
    Item it = item_of(geo, item0);
    for (int j = 0; j < count; ++j, advance(geo, it)) {
      const int rows_here = it.c_rows + it.p_rows;
      mbar_wait(BAR(B_PFULL), (uint32_t)(j & 1));  // the row sums are in shared memory
      mbar_wait(BAR(B_OFULL), (uint32_t)(j & 1));
      tc_fence_after();
      if (tracer) ATTN_TRACE(j, 7);
      if (q * 32 < rows_here) {
        const uint32_t rsum = red_sum + (uint32_t)(((j & 1) * PARTS * QT + row) * 4);
        const float inv = 1.0f / ((lds_f32(rsum) + lds_f32(rsum + QT * 4)) + (lds_f32(rsum + 2 * QT * 4) + lds_f32(rsum + 3 * QT * 4)));
        T *dst = nullptr;
        if (row < it.c_rows)
          dst = out_ctx + ((long long)it.g * n + it.t * QT + row) * D + it.h * HD;
        else if (row < rows_here)
          dst = out_prompt + ((long long)it.g * K + (row - it.c_rows)) * D + it.h * HD;
#pragma unroll
        for (int c = 0; c < 4; ++c) {  // 16 columns at a time: this role runs on 40 registers
          uint32_t acc[16];
          tmem_ld16(taddr + (uint32_t)(c * 16), acc);
          if (dst) {
#pragma unroll
            for (int v = 0; v < 2; ++v) {
              uint4 o;
              o.x = pack2<T>(__uint_as_float(acc[8 * v + 0]) * inv, __uint_as_float(acc[8 * v + 1]) * inv);
              o.y = pack2<T>(__uint_as_float(acc[8 * v + 2]) * inv, __uint_as_float(acc[8 * v + 3]) * inv);
              o.z = pack2<T>(__uint_as_float(acc[8 * v + 4]) * inv, __uint_as_float(acc[8 * v + 5]) * inv);
              o.w = pack2<T>(__uint_as_float(acc[8 * v + 6]) * inv, __uint_as_float(acc[8 * v + 7]) * inv);
              *reinterpret_cast<uint4 *>(dst + c * 16 + v * 8) = o;
            }
          }
        }
      }
      // O is out of tensor memory: the next P V may overwrite it
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_OFREE));
      if (tracer) ATTN_TRACE(j, 8);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) ATTN_TRACE(0, 14);
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (threadIdx.x == 0) ATTN_TRACE(0, 15);
}

// ================================================================================================================
// Two-slot form (n16 <= 224: ViT-B/16).  The softmax of consecutive tiles is taken by two groups of eight warps in
// turn (tile j -> group j % 2 -> TMEM slot j % 2), TWO THREADS PER QUERY ROW, each with one half of the row's 16-key
// blocks.  A thread first takes the maximum of its half (four 16-column tensor-memory loads in flight) and swaps it
// with the other half's through shared memory (named barrier of the two warps), then makes the second pass: scores
// out of tensor memory 16 at a time (the next 16 in flight), p = ex2(s / 8 log2 e - max / 8 log2 e) in packed f32x2
// arithmetic, f32 row sum, and the rounded probabilities go back into tensor memory over the thread's OWN scores
// (two 16-bit values per 32-bit column: the layout tcgen05.mma reads an A operand from), starting at the first column
// of its half -- the two threads of a row never touch each other's columns.  O = P V takes A straight from tensor
// memory, one K = 16 instruction per key block.  While one group waits for its next tile (P V of the slot's previous
// tile, then its S) the other group's exponentials have the MUFU pipe (16 ex2 / clk / SM) to themselves.
//
// Measured on B200 (tools/attn_trace.py, tools/probe/): MUFU.EX2 issues once per 8 clocks per scheduler whatever
// else runs (F2FP, FFMA2, FADD2 ride beside it); tensor-memory loads are not the limit (> 860 B/clk/SM); what the
// 197-key tile costs beyond its 1664 MUFU clocks is the dependent chain P stored -> P V -> S of the slot's next
// tile -> maximum, about 1.8k clocks in which only one group has work.  Rejected after measurement: the maximum on
// the epilogue warps (their wake-up sits on that chain), one thread per row (a single warp per scheduler cannot
// keep the MUFU pipe fed: 200 clocks per key block instead of 131), S(j+2) issued by the P V thread straight behind
// P V(j) (completes ~1000 clocks later than from its own thread after the commit), a share of the exponentials as a
// degree-4 polynomial on the FMA pipe (2 / 3 / 4 / 5 of every 8 key pairs: 17.2 / 17.6 / 18.2 / 18.9 us against 17.2:
// the softmax pass is bound by instruction issue and latency, not by the MUFU pipe alone).
//
//   warp 0 (one thread)   TMA producer (as above)
//   warp 1 (one thread)   S = Q K^T into slot j % 2 once P V of tile j-2 has consumed that slot's P
//   warp 2 (one thread)   O = P V (A = P in TMEM, B = V as loaded, MN-major) into the 64 columns behind the slots
//   warps 4..11 / 12..19  softmax groups 0 / 1: warp = (half of the keys, lane quarter)
//   warps 20..23          epilogue: O out of TMEM, times 1 / row sum, into a 128B-swizzled staging tile that one
//                         thread hands to the TMA (cp.async.bulk.tensor store: context rows and prompt rows are two
//                         boxes of the tile)
// Registers (setmaxnreg, out of 768 x 80): producers 24, softmax 96, epilogue 72.
static constexpr int PP_THREADS = 768;
static constexpr int PP_SLOT_COLS = 224;
static constexpr int PP_O_COL = 448;
static constexpr int PP_Q_RING = 4;
static constexpr int PP_O_STAGES = 3;
static constexpr int PP_REGS_PROD = 24, PP_REGS_SOFTMAX = 96, PP_REGS_EPI = 72;  // out of 768 x 80
static constexpr int PP_AUX_WARP0 = 20;

// N_CT: number of keys when known at compile time (197: ViT-B/16 -- the softmax passes are then straight-line code the
// scheduler can pipeline across key blocks, and the padding keys of the last block cost no exponentials), 0: taken
// from geo at run time
template <typename T, int N_CT>
__global__ void __launch_bounds__(PP_THREADS, 1)
    ro_attn_fwd_pp(const __grid_constant__ CUtensorMap map_full,   // q|k|v matrix, box 64 x 128
                   const __grid_constant__ CUtensorMap map_kvt,    // q|k|v matrix, box 64 x (n16 % 128)
                   const __grid_constant__ CUtensorMap map_qt,     // q|k|v matrix, box 64 x (n % 128)
                   const __grid_constant__ CUtensorMap map_prompt, // prompt-q matrix, box 64 x K
                   const __grid_constant__ CUtensorMap omap_full,   // context output, box 64 x 128
                   const __grid_constant__ CUtensorMap omap_tail,   // context output, box 64 x (n % 128)
                   const __grid_constant__ CUtensorMap omap_prompt, // prompt output, box 64 x K
                   Geo geo, int total_items) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int NBLK_CT = (N_CT + 15) / 16;
  const int n = N_CT ? N_CT : geo.n, n16 = geo.n16, K = geo.K, H = geo.H;
  const int D = H * HD;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv_bytes = n16 * ROW_BYTES;
  // shared memory: [K|V ring of 3 | Q ring | O staging x 3 | barriers | row sums | row maxima]
  const uint32_t kv_base = smem_u32(smem);
  const uint32_t q_base = kv_base + KV_RING * kv_bytes;
  const uint32_t o_base = q_base + PP_Q_RING * Q_TILE_BYTES;
  uint8_t *tail = smem + KV_RING * kv_bytes + (PP_Q_RING + PP_O_STAGES) * Q_TILE_BYTES;
  uint64_t *bars = reinterpret_cast<uint64_t *>(tail);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int idx) { return bar0 + 8u * idx; };
  constexpr int B_KVFULL = 0, B_KVFREE = 3, B_QFULL = 6, B_QFREE = 10, B_SFULL = 14, B_SFREE = 16, B_PFULL = 18,
                B_OFULL = 20, B_OFREE = 21, N_BARS = 22;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + N_BARS);
  // f32 [4][2][QT]: the two partial row sums of tile j in buffer j % 4 (written at the end of the softmax of tile j, read by the
  // epilogue of tile j; the softmax of tile j+4 starts after P V of tile j+2, which waited for the epilogue of j+1)
  const uint32_t red_sum = smem_u32(bars + N_BARS + 1);
  // f32 [2][2][QT]: the maxima the two threads of a row (group, half) found in their halves of the keys
  const uint32_t row_max = red_sum + 4 * 2 * QT * 4;

  const int item0 = (int)((long long)total_items * blockIdx.x / gridDim.x);
  const int item1 = (int)((long long)total_items * (blockIdx.x + 1) / gridDim.x);
  const int count = item1 - item0;
#ifdef RPO_DIAG
  long long *const trace_buf = g_attn_trace;
  if (threadIdx.x == 0) {
    ATTN_TRACE(0, 12);
    ATTN_WALL(0);
  }
#endif

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_full)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_kvt)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_qt)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_prompt)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&omap_full)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&omap_tail)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&omap_prompt)) : "memory");
    for (int i = 0; i < KV_RING; ++i) {
      mbar_init(BAR(B_KVFULL + i), 1);
      mbar_init(BAR(B_KVFREE + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(BAR(B_SFULL + i), 1);
      mbar_init(BAR(B_SFREE + i), 1);
      mbar_init(BAR(B_PFULL + i), 8);
    }
    mbar_init(BAR(B_OFULL), 1);
    mbar_init(BAR(B_OFREE), 4);
    for (int i = 0; i < PP_Q_RING; ++i) {
      mbar_init(BAR(B_QFULL + i), 1);
      mbar_init(BAR(B_QFREE + i), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  if (threadIdx.x == 0) ATTN_TRACE(0, 13);

  // K of unit u (= image, head; counted from the CTA's first) lives in ring entry (2u) % 3, V in (2u + 1) % 3
  const int tiles = geo.tiles;
  const int unit0 = item0 / tiles;
  auto unit_idx = [&](int j) { return (item0 + j) / tiles - unit0; };
  auto first_of_unit = [&](int j) { return j == 0 || (item0 + j) % tiles == 0; };
  auto last_of_unit = [&](int j) { return j + 1 == count || (item0 + j + 1) % tiles == 0; };

  if (warp < 4) setmaxnreg_dec<PP_REGS_PROD>();
  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      pdl_wait();
      ATTN_WALL(1);
      for (int j = 0; j < count; ++j) {
        const Item it = item_of(geo, item0 + j);
        if (first_of_unit(j)) {
          const int u = unit_idx(j);
#pragma unroll
          for (int kv = 0; kv < 2; ++kv) {  // K, then V
            const int x = 2 * u + kv, e = x % KV_RING, use = x / KV_RING;
            if (use >= 1) mbar_wait(BAR(B_KVFREE + e), (uint32_t)((use - 1) & 1));
            const uint32_t dst = kv_base + e * kv_bytes;
            mbar_arrive_expect_tx(BAR(B_KVFULL + e), (uint32_t)kv_bytes);
            for (int r = 0; r < n16; r += 128)
              tma_load_2d(dst + r * ROW_BYTES, (n16 - r >= 128) ? &map_full : &map_kvt, BAR(B_KVFULL + e),
                          (1 + kv) * D + it.h * HD, it.g * n + r);
          }
        }
        const int qs = j % PP_Q_RING;
        if (j >= PP_Q_RING) mbar_wait(BAR(B_QFREE + qs), (uint32_t)(((j / PP_Q_RING) - 1) & 1));
        const uint32_t Qs = q_base + qs * Q_TILE_BYTES;
        mbar_arrive_expect_tx(BAR(B_QFULL + qs), (uint32_t)((it.c_rows + it.p_rows) * ROW_BYTES));
        if (it.c_rows == QT)
          tma_load_2d(Qs, &map_full, BAR(B_QFULL + qs), it.h * HD, it.g * n + it.t * QT);
        else if (it.c_rows > 0)
          tma_load_2d(Qs, &map_qt, BAR(B_QFULL + qs), it.h * HD, it.g * n + it.t * QT);
        if (it.p_rows > 0)
          tma_load_2d(Qs + geo.prompt_row * ROW_BYTES, &map_prompt, BAR(B_QFULL + qs), it.h * HD, it.g * K);
        ATTN_TRACE(j, 10);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ===== S = Q K^T issuer =====
      const uint32_t fmt = Num<T>::dtype == RPO_BF16 ? 1u : 0u;
      const uint32_t idesc_s = make_idesc((int)fmt, QT, n16);
      for (int j = 0; j < count; ++j) {
        const int slot = j & 1, qs = j % PP_Q_RING;
        const int x = 2 * unit_idx(j), e = x % KV_RING;
        mbar_wait_spin(BAR(B_QFULL + qs), (uint32_t)((j / PP_Q_RING) & 1));
        if (first_of_unit(j)) mbar_wait_spin(BAR(B_KVFULL + e), (uint32_t)((x / KV_RING) & 1));
        // the slot's previous P has been consumed by its P V.  (Issuing this S right behind that P V from ONE thread,
        // relying on the tensor core's issue order instead of this wait, was measured: the S then completes ~1000
        // clocks later than it does this way.)
        if (j >= 2) mbar_wait_spin(BAR(B_SFREE + slot), (uint32_t)(((j >> 1) - 1) & 1));
        tc_fence_after();
        const uint32_t tslot = tmem_base + (uint32_t)(slot * PP_SLOT_COLS);
        const uint64_t adesc = make_smem_desc(q_base + qs * Q_TILE_BYTES), bdesc = make_smem_desc(kv_base + e * kv_bytes);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_f16(tslot, adesc + 2u * k, bdesc + 2u * k, idesc_s, k != 0);
        umma_commit(BAR(B_SFULL + slot));
        umma_commit(BAR(B_QFREE + qs));
        if (last_of_unit(j)) umma_commit(BAR(B_KVFREE + e));  // K of the unit is dead once this S is complete
        ATTN_TRACE(j, 0);
      }
    }
  } else if (warp == 2) {
    if (elect_one()) {
      // ===== O = P V issuer =====
      const uint32_t fmt = Num<T>::dtype == RPO_BF16 ? 1u : 0u;
      const uint32_t idesc_o = make_idesc((int)fmt, QT, HD) | (1u << 16);  // B (= V) is MN-major
      const int nblk = NBLK_CT ? NBLK_CT : geo.nblk;
      const int split = (nblk + 1) >> 1;
      for (int j = 0; j < count; ++j) {
        const int slot = j & 1;
        const int x = 2 * unit_idx(j) + 1, e = x % KV_RING;
        if (first_of_unit(j)) mbar_wait_spin(BAR(B_KVFULL + e), (uint32_t)((x / KV_RING) & 1));
        if (j >= 1) mbar_wait_spin(BAR(B_OFREE), (uint32_t)((j - 1) & 1));  // the previous tile's O is out of tensor memory
        mbar_wait_spin(BAR(B_PFULL + slot), (uint32_t)((j >> 1) & 1));
        ATTN_TRACE(j, 1);
        tc_fence_after();
        const uint64_t vdesc = make_smem_desc(kv_base + e * kv_bytes);
        const uint32_t pcol = tmem_base + (uint32_t)(slot * PP_SLOT_COLS);
        // 16 keys per instruction: 8 columns of P, 2 KB of V.  The P of the keys of each softmax thread starts where
        // that thread's scores started: key blocks [0, split) at column 0, [split, nblk) at column 16 * split.
        for (int blk = 0; blk < nblk; ++blk)
          umma_f16_ts(tmem_base + PP_O_COL, pcol + (uint32_t)(blk < split ? blk * 8 : split * 8 + blk * 8),
                      vdesc + (uint64_t)(blk * 128), idesc_o, blk != 0);
        umma_commit(BAR(B_OFULL));
        umma_commit(BAR(B_SFREE + slot));
        if (last_of_unit(j)) umma_commit(BAR(B_KVFREE + e));  // V of the unit is dead once this O is complete
        ATTN_TRACE(j, 2);
      }
    }
  } else if (warp >= 4 && warp < PP_AUX_WARP0) {
    // ===== softmax: group wg takes tiles j = wg, wg + 2, ...; two threads per query row (= TMEM lane), one per half
    // of the key blocks.  Pass 1: maximum over the thread's half (four 16-key blocks in flight), exchanged with the
    // other half through shared memory (named barrier of the two warps).  Pass 2: p = ex2(s / 8 log2 e - max ...),
    // f32 sum, and the rounded probabilities over the thread's OWN scores (from the first column of its half), so the
    // two threads of a row never touch each other's columns. =====
    setmaxnreg_inc<PP_REGS_SOFTMAX>();
    const int wg = (warp - 4) >> 3;
    const int half = ((warp - 4) >> 2) & 1;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool tracer = ((threadIdx.x - 128) & 255) == 0;
    const float sl2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    const int nblk = NBLK_CT ? NBLK_CT : geo.nblk;
    const int split = (nblk + 1) >> 1;
    const int pad_first = n - (nblk - 1) * 16;  // valid keys of the last 16-key block
    // this thread's scores, and (over them) its probabilities
    const uint32_t sbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(wg * PP_SLOT_COLS + half * split * 16);
    const uint32_t my_max = row_max + (uint32_t)(((wg * 2 + half) * QT + row) * 4);
    const uint32_t other_max = row_max + (uint32_t)(((wg * 2 + (half ^ 1)) * QT + row) * 4);
    const int pair_bar = 2 + wg * 4 + q;  // named barrier of the two warps that share this lane quarter
    auto mask_pad = [&](uint32_t (&x)[16]) {
#pragma unroll
      for (int e = 0; e < 16; ++e)
        if (e >= pad_first) x[e] = 0xff800000u;  // -inf: drops out of the maximum, ex2 gives 0
    };
    // cnt = key blocks of this thread (compile-time when the kernel knows the key count); has_pad: its last block
    // is the row's last
    auto tile = [&](auto cnt_c, int cnt_rt, bool has_pad, int j) {
      constexpr int CNT_CT = decltype(cnt_c)::value;
      constexpr int CNT_MAX = CNT_CT ? CNT_CT : 7;
      const int cnt = CNT_CT ? CNT_CT : cnt_rt;
      uint32_t a[16], b[16];
      float mx;
      {
        // ---- pass 1 ----
        uint32_t c[16], d[16];
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        auto fold = [&](bool last, uint32_t (&x)[16]) {
          if (last) mask_pad(x);
#pragma unroll
          for (int e = 0; e < 16; e += 4)
            m4[e >> 2] = fmaxf(m4[e >> 2], fmaxf(fmaxf(__uint_as_float(x[e]), __uint_as_float(x[e + 1])),
                                                 fmaxf(__uint_as_float(x[e + 2]), __uint_as_float(x[e + 3]))));
        };
#pragma unroll
        for (int b0 = 0; b0 < CNT_MAX; b0 += 4) {
          if (b0 < cnt) {
            // unconditional loads: columns behind the thread's last block are somebody else's, and never folded
            tmem_ld16_nowait(sbase + (uint32_t)(b0 * 16), a);
            tmem_ld16_nowait(sbase + (uint32_t)((b0 + 1) * 16), b);
            tmem_ld16_nowait(sbase + (uint32_t)((b0 + 2) * 16), c);
            tmem_ld16_nowait(sbase + (uint32_t)((b0 + 3) * 16), d);
            tmem_ld_wait(a);
            tmem_ld_wait(b);
            tmem_ld_wait(c);
            tmem_ld_wait(d);
            if (tracer && b0 == 0) ATTN_TRACE(j, 4);
            fold(has_pad && b0 == cnt - 1, a);
            if (b0 + 1 < cnt) fold(has_pad && b0 + 1 == cnt - 1, b);
            if (b0 + 2 < cnt) fold(has_pad && b0 + 2 == cnt - 1, c);
            if (b0 + 3 < cnt) fold(has_pad && b0 + 3 == cnt - 1, d);
          }
        }
        mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));  // -inf if the thread has no block
      }
      if (tracer) ATTN_TRACE(j, 9);
      if (cnt > 0) tmem_ld16_nowait(sbase, a);  // pass 2's first block, behind the exchange
      sts_f32(my_max, mx);
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      const float off = fmaxf(mx, lds_f32(other_max)) * sl2;  // finite: every row sees key 0
      if (tracer) ATTN_TRACE(j, 5);
      // ---- pass 2 ----
      const uint64_t sl2x2 = pack_f32x2(sl2, sl2), noff2 = pack_f32x2(-off, -off);
      uint64_t lsum = pack_f32x2(0.f, 0.f);
      auto prob_block = [&](int i, bool last, uint32_t (&sc)[16]) {
        if (last) mask_pad(sc);
        uint32_t pk[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          if (last && 2 * e >= pad_first) {  // two padding keys: p = 0 without the exponentials
            pk[e] = 0u;
            continue;
          }
          float x0, x1;
          unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(sc[2 * e]), __uint_as_float(sc[2 * e + 1])), sl2x2, noff2), x0, x1);
          const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
          lsum = add_f32x2(lsum, pack_f32x2(p0, p1));  // even keys in the low lane, odd keys in the high lane
          pk[e] = pack2<T>(p0, p1);
        }
        tmem_st8(sbase + (uint32_t)(i * 8), pk);
      };
#pragma unroll
      for (int i = 0; i < CNT_MAX; i += 2) {
        if (i < cnt) {
          tmem_ld_wait(a);
          tmem_ld16_nowait(sbase + (uint32_t)((i + 1) * 16), b);  // (behind the thread's last block: never used)
          prob_block(i, has_pad && i == cnt - 1, a);
          if (i + 1 < cnt) {
            tmem_ld_wait(b);
            tmem_ld16_nowait(sbase + (uint32_t)((i + 2) * 16), a);
            prob_block(i + 1, has_pad && i + 1 == cnt - 1, b);
          }
          if (tracer && i == 0) ATTN_TRACE(j, 11);
          if (tracer && i == 2 && j > 0) ATTN_TRACE(j, 12);
          if (tracer && i == 4 && j > 0) ATTN_TRACE(j, 13);
        }
      }
      if (tracer && j > 0) ATTN_TRACE(j, 14);
      tmem_ld_wait(a);  // no load is left in flight
      tmem_ld_wait(b);
      float l0, l1;
      unpack_f32x2(lsum, l0, l1);
      sts_f32(red_sum + (uint32_t)((((j & 3) * 2 + half) * QT + row) * 4), l0 + l1);
      tmem_st_wait();
      if (tracer && j > 0) ATTN_TRACE(j, 15);
    };
    for (int j = wg; j < count; j += 2) {
      const Item it = item_of(geo, item0 + j);
      const bool valid = q * 32 < it.c_rows + it.p_rows;  // warp-uniform, the same for the two warps of a quarter
      mbar_wait(BAR(B_SFULL + wg), (uint32_t)((j >> 1) & 1));
      if (tracer) ATTN_TRACE(j, 3);
      if (valid) {
        tc_fence_after();
        if (NBLK_CT) {
          constexpr int SPLIT = (NBLK_CT + 1) / 2;
          if (half == 0)
            tile(std::integral_constant<int, SPLIT>{}, 0, SPLIT == NBLK_CT, j);
          else if (NBLK_CT > SPLIT)
            tile(std::integral_constant<int, (NBLK_CT > SPLIT ? NBLK_CT - SPLIT : 1)>{}, 0, true, j);
          else
            tile(std::integral_constant<int, 0>{}, 0, false, j);
        } else {
          tile(std::integral_constant<int, 0>{}, half == 0 ? split : nblk - split, half == 0 ? split == nblk : true, j);
        }
        tc_fence_before();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_PFULL + wg));
      if (tracer) ATTN_TRACE(j, 6);
    }
  } else if (warp >= PP_AUX_WARP0) {
    // ===== epilogue: one thread per query row =====
    setmaxnreg_dec<PP_REGS_EPI>();
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool tracer = threadIdx.x == PP_AUX_WARP0 * 32;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + PP_O_COL;
    pdl_wait();
    Item it = item_of(geo, item0);
    for (int j = 0; j < count; ++j, advance(geo, it)) {
      const int rows_here = it.c_rows + it.p_rows;
      mbar_wait(BAR(B_OFULL), (uint32_t)(j & 1));
      tc_fence_after();
      if (tracer) ATTN_TRACE(j, 7);
      const bool valid = q * 32 < rows_here;
      // O times 1 / row sum into the staging tile (rows of 128 B, 16-byte chunk index ^ (row & 7): the layout a
      // 128B-swizzled tensor map reads), then ONE thread hands the tile to the TMA: context rows and prompt rows are
      // two boxes of it.  Three staging tiles: the tile written now was read by the store of three tiles ago, which
      // the issuing thread has waited for before the previous tile's barrier.
      const uint32_t stage = o_base + (uint32_t)((j % PP_O_STAGES) * Q_TILE_BYTES);
      if (valid) {
        const uint32_t rs = red_sum + (uint32_t)((((j & 3) * 2) * QT + row) * 4);
        const float inv = 1.0f / (lds_f32(rs) + lds_f32(rs + QT * 4));
        const uint32_t srow = stage + (uint32_t)(row * ROW_BYTES);
        const uint32_t sw = (uint32_t)(row & 7);
#pragma unroll
        for (int c = 0; c < 4; c += 2) {
          uint32_t acc[2][16];
          tmem_ld16_nowait(taddr + (uint32_t)(c * 16), acc[0]);
          tmem_ld16_nowait(taddr + (uint32_t)((c + 1) * 16), acc[1]);
          tmem_ld_wait(acc[0]);
          tmem_ld_wait(acc[1]);
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
#pragma unroll
            for (int v = 0; v < 2; ++v) {
              sts128(srow + ((((uint32_t)(2 * (c + cc) + v)) ^ sw) << 4),
                     pack2<T>(__uint_as_float(acc[cc][8 * v + 0]) * inv, __uint_as_float(acc[cc][8 * v + 1]) * inv),
                     pack2<T>(__uint_as_float(acc[cc][8 * v + 2]) * inv, __uint_as_float(acc[cc][8 * v + 3]) * inv),
                     pack2<T>(__uint_as_float(acc[cc][8 * v + 4]) * inv, __uint_as_float(acc[cc][8 * v + 5]) * inv),
                     pack2<T>(__uint_as_float(acc[cc][8 * v + 6]) * inv, __uint_as_float(acc[cc][8 * v + 7]) * inv));
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
      // O is out of tensor memory: the next P V may overwrite it
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(B_OFREE));
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == PP_AUX_WARP0 * 32) {
        if (it.c_rows == QT)
          tma_store_2d(&omap_full, stage, it.h * HD, it.g * n + it.t * QT);
        else if (it.c_rows > 0)
          tma_store_2d(&omap_tail, stage, it.h * HD, it.g * n + it.t * QT);
        if (it.p_rows > 0)
          tma_store_2d(&omap_prompt, stage + (uint32_t)(geo.prompt_row * ROW_BYTES), it.h * HD, it.g * K);
        tma_store_commit();
        tma_store_wait_read<1>();  // the store of the previous tile has read its staging tile
      }
      if (tracer) ATTN_TRACE(j, 8);
    }
    if (threadIdx.x == PP_AUX_WARP0 * 32) tma_store_wait_read<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    ATTN_TRACE(0, 14);
    ATTN_WALL(2);
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (threadIdx.x == 0) {
    ATTN_TRACE(0, 15);
    ATTN_WALL(3);
  }
}

static int pp_smem_bytes(int n16) {
  return KV_RING * n16 * ROW_BYTES + (PP_Q_RING + PP_O_STAGES) * Q_TILE_BYTES + 8 * 24 + (8 + 4) * QT * 4 + 1024;
}

static int smem_bytes(int n16, int q_ring) {
  return KV_RING * n16 * ROW_BYTES + (((n16 >> 4) + 3) / 4) * P_TILE_BYTES + q_ring * Q_TILE_BYTES + 8 * 28 +
         (2 + 2) * PARTS * QT * 4 + 1024;
}

}  // namespace atc

bool ro_attention_fwd_dense_supported(int dtype, int n, int K, int H) {
  if (dtype != RPO_F16 && dtype != RPO_BF16) return false;
  if (n < 1 || K < 0 || H < 1) return false;
  const int n16 = (n + 15) & ~15;
  if (n16 > 288) return false;                     // 18 key blocks: 5 per softmax thread at most; S + O within 512 TMEM columns (shared memory stops at 272)
  if (K > 0 && (n % 128) + K > 128) return false;  // all prompt rows in one query tile
  if ((n16 <= atc::PP_SLOT_COLS ? atc::pp_smem_bytes(n16) : atc::smem_bytes(n16, 2)) > 227 * 1024) return false;
  return true;
}

template <typename T>
int ro_attention_fwd_dense(const T *qkv_ctx, const T *q_prompt, T *out_ctx, T *out_prompt, int G, int n, int K, int H,
                           cudaStream_t st) {
  if constexpr (sizeof(T) != 2) {
    set_error("tcgen05 attention supports f16/bf16 only");
    return RPO_ERR_INVALID;
  } else {
    using namespace atc;
    RPO_REQUIRE(ro_attention_fwd_dense_supported(Num<T>::dtype, n, K, H), "tcgen05 attention shape");
    RPO_REQUIRE(G >= 1, "at least one group");
    RPO_REQUIRE((((uintptr_t)qkv_ctx | (uintptr_t)q_prompt | (uintptr_t)out_ctx | (uintptr_t)out_prompt) & 15) == 0,
                "attention buffers must be 16-byte aligned");
    const int D = H * HD;
    const int n16 = (n + 15) & ~15;
    Geo geo;
    geo.n = n;
    geo.n16 = n16;
    geo.nblk = n16 >> 4;
    geo.K = K;
    geo.prompt_tile = n / 128;
    geo.prompt_row = n % 128;
    geo.H = H;
    geo.tiles = K > 0 ? geo.prompt_tile + 1 : (n + 127) / 128;
    // S as one UMMA N block up to 256 keys, else two (144 + the rest: both multiples of 16, the second starts on an
    // 8-row group of the 128B-swizzled K tile)
    geo.n_first = n16 <= 256 ? n16 : 144;
    const bool two_slots = n16 <= PP_SLOT_COLS;  // two S / P slots of 224 columns beside the 64 columns of O
    int q_ring = MAX_Q_RING;
    while (q_ring > 2 && smem_bytes(n16, q_ring) > 227 * 1024) --q_ring;
    geo.q_ring = q_ring;
    const int smem = two_slots ? pp_smem_bytes(n16) : smem_bytes(n16, q_ring);
    CUtensorMap map_full, map_kvt, map_qt, map_prompt;
    const long long Mc = (long long)G * n;
    RPO_TRY(make_map(&map_full, Num<T>::dtype, qkv_ctx, Mc, 3 * D, 3LL * D, 128));
    RPO_TRY(make_map(&map_kvt, Num<T>::dtype, qkv_ctx, Mc, 3 * D, 3LL * D, n16 % 128 ? n16 % 128 : 128));
    RPO_TRY(make_map(&map_qt, Num<T>::dtype, qkv_ctx, Mc, 3 * D, 3LL * D, n % 128 ? n % 128 : 128));
    if (K > 0)
      RPO_TRY(make_map(&map_prompt, Num<T>::dtype, q_prompt, (long long)G * K, D, D, K));
    else
      map_prompt = map_full;
    const long long items = (long long)geo.tiles * H * G;
    RPO_REQUIRE(items <= 0x7fffffffLL, "grid limits");
    const int sms = sm_count();
    dim3 grid((unsigned)(items < sms ? items : sms));
    prof_tag("attn_fwd_tc G=%d H=%d K=%d n=%d", G, H, K, n);
    // all instantiations of a kernel share one function-pointer type: remember the configured size per kernel
    static std::map<const void *, int> configured;
    auto configure = [&](const void *kernel) -> int {
      int &have = configured[kernel];
      if (smem > have) {
        RPO_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        have = smem;
      }
      return RPO_OK;
    };
    int s = RPO_ERR_INVALID;
    if (two_slots) {
      // the output leaves through the TMA as well: context tiles (full / last) and the prompt rows
      CUtensorMap omap_full, omap_tail, omap_prompt;
      RPO_TRY(make_map(&omap_full, Num<T>::dtype, out_ctx, Mc, D, D, 128));
      RPO_TRY(make_map(&omap_tail, Num<T>::dtype, out_ctx, Mc, D, D, n % 128 ? n % 128 : 128));
      if (K > 0)
        RPO_TRY(make_map(&omap_prompt, Num<T>::dtype, out_prompt, (long long)G * K, D, D, K));
      else
        omap_prompt = omap_full;
      auto launch = [&](auto kernel) -> int {
        RPO_TRY(configure(reinterpret_cast<const void *>(kernel)));
        RPO_CHECK_CUDA(launch_pdl(kernel, grid, dim3(PP_THREADS), (size_t)smem, st, map_full, map_kvt, map_qt, map_prompt,
                                  omap_full, omap_tail, omap_prompt, geo, (int)items));
        return RPO_OK;
      };
      s = n == 197 ? launch(ro_attn_fwd_pp<T, 197>) : launch(ro_attn_fwd_pp<T, 0>);
    } else {
      auto launch = [&](auto kernel) -> int {
        RPO_TRY(configure(reinterpret_cast<const void *>(kernel)));
        RPO_CHECK_CUDA(launch_pdl(kernel, grid, dim3(THREADS), (size_t)smem, st, map_full, map_kvt, map_qt, map_prompt,
                                  out_ctx, out_prompt, geo, (int)items));
        return RPO_OK;
      };
      // 15 .. 18 key blocks: MAXB = key blocks of the busiest softmax thread (see the kernel)
      const int rounds = geo.nblk / PARTS + (geo.nblk % PARTS ? 1 : 0);
      s = rounds <= 4 ? launch(ro_attn_fwd_tc<T, 4>) : launch(ro_attn_fwd_tc<T, 5>);
    }
    RPO_TRY(s);
    RPO_LAUNCH_CHECK();
    return RPO_OK;
  }
}

#ifdef RPO_DIAG
extern "C" int rpo_diag_set_attn_trace(long long *buf) {
  RPO_CHECK_CUDA(cudaMemcpyToSymbol(atc::g_attn_trace, &buf, sizeof(buf)));
  return RPO_OK;
}
#endif

template int ro_attention_fwd_dense<float>(const float *, const float *, float *, float *, int, int, int, int,
                                           cudaStream_t);
template int ro_attention_fwd_dense<__half>(const __half *, const __half *, __half *, __half *, int, int, int, int,
                                            cudaStream_t);
template int ro_attention_fwd_dense<__nv_bfloat16>(const __nv_bfloat16 *, const __nv_bfloat16 *, __nv_bfloat16 *,
                                                   __nv_bfloat16 *, int, int, int, int, cudaStream_t);

}  // namespace rpo
