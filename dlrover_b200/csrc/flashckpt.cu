// flashckpt.cu — device engine of the B200 Flash Checkpoint path (sm_100a only).
//
// What the reference does on this path (dlrover @ 468d632):
//   ckpt_saver.py:198-231   per leaf: torch.frombuffer(shm).copy_(gpu_tensor)
//                           -> one blocking pageable cudaMemcpy per tensor.
// What this file does instead:
//   1. a cached descriptor table ("plan") maps every leaf tensor to its byte
//      offset in the checkpoint segment (reference layout: running sum of
//      numel*element_size, no padding; ckpt_saver.py:286-301);
//   2. ONE persistent gather kernel copies all leaves into a contiguous HBM
//      arena that is a byte image of the segment (pack), HBM-bandwidth bound:
//      algorithmic traffic 2*S bytes per save;
//   3. the arena is drained to the (pinned, NUMA-local) POSIX shm segment by DMA
//      on a side stream gated by an event — the training stream only ever waits
//      for (2).  A pump thread feeds the copy engine ONE 32 MiB piece at a time
//      so the drain does not starve the application's own D2H copies;
//   4. restore is the inverse: DMA fill + scatter kernel;
//   5. if the state does not fit in HBM a second time, (2)+(3) run window by
//      window through a bounded arena (blocking, PCIe speed).
//
// Three kernels, all templated on direction (0 = gather/pack, 1 = scatter):
//   fc_copy_tma       one elected thread per CTA drives a ring of cp.async.bulk
//                     global->shared (mbarrier complete_tx) and shared->global
//                     (bulk_group) transfers: the 16-B congruent bodies.
//                     Default 148 CTAs x 2 stages x 96 KiB: 4.76 ms for 32 GB of
//                     traffic, 1.02x the measured copy peak.
//   fc_copy_tma_shift ranges whose source and destination are NOT congruent
//                     mod 16: TMA in, funnel shift shared->shared, TMA out.
//   fc_copy_lsu       256-thread CTAs, 128-bit LDG/STG, 4-way unrolled: heads,
//                     tails, tiny ranges, the bounded-arena windows, and
//                     everything when FC_VARIANT_LSU is selected.
//
// No torch types here; see include/flashckpt.h for the ABI contract.

#include "../../include/flashckpt.h"

#include <cuda_runtime.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

// ------------------------------------------------------------------ errors --

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
  snprintf(g_err, sizeof(g_err), fmt, a, b);
  return code;
}

#define FC_CUDA(call)                                                          \
  do {                                                                         \
    cudaError_t _e = (call);                                                   \
    if (_e != cudaSuccess) return fail(FC_ECUDA, "%s: %s", #call, cudaGetErrorString(_e)); \
  } while (0)

extern "C" int fc_version(void) { return FC_VERSION; }

extern "C" const char* fc_strerror(int code) {
  switch (code) {
    case FC_OK: return "ok";
    case FC_EINVAL: return "invalid argument";
    case FC_ECUDA: return "CUDA runtime error";
    case FC_ENOMEM: return "out of memory";
    case FC_EBUSY: return "a save/restore is still in flight";
    case FC_ENOTREADY: return "not ready";
    default: return "unknown flashckpt error";
  }
}

extern "C" const char* fc_last_error(void) { return g_err; }

// --------------------------------------------------------------- work items --

// One work item = one contiguous byte range of one tensor, <= chunk_bytes.
// 32 bytes so a CTA fetches it with two 16-B loads.
struct __align__(16) FcItem {
  uint64_t tptr;    // device address inside the tensor
  uint64_t aoff;    // byte offset inside the arena (== offset in the shm segment)
  uint32_t nbytes;  // > 0
  uint32_t pad0;
  uint64_t pad1;
};
static_assert(sizeof(FcItem) == 32, "FcItem must be 32 bytes");

struct FcRun {  // merged contiguous arena range (drain/fill DMA granularity)
  uint64_t off;
  uint64_t len;
};

struct FcSpan {  // one input range (a tensor's bytes), ascending arena offset
  uint64_t tptr;
  uint64_t off;
  uint64_t len;
};

constexpr int kLsuThreads = 256;
constexpr int kLsuUnroll = 4;
constexpr uint32_t kDefaultChunk = 256u << 10;      // work-item size
constexpr uint64_t kDmaPiece = 256ull << 20;        // restore fill memcpy size
// Drain pacing.  Measured on B200 (tools/d2h_probe2.py, profiles/r01_d2h_pacing.md):
// while a stream has ANOTHER D2H copy queued behind the one in flight, the
// copy engine keeps serving that stream, and a small D2H copy from any other
// stream of the process (a `loss.item()`!) starves until the whole checkpoint
// has left the device (145-290 ms).  With exactly ONE piece in flight the
// engine's queue empties for a moment after every piece and the foreign copy
// goes through in ~0.4 ms, at 54.5 instead of 55.2 GB/s of drain throughput.
// The pump thread therefore submits piece k+1 only after piece k completed.
constexpr uint64_t kDrainPiece = 32ull << 20;
constexpr int kDrainDepth = 1;
constexpr int kDrainRing = 8;

// ------------------------------------------------------------ device helpers --

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

__device__ __forceinline__ uint4 ldg_cached(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

__device__ __forceinline__ void stg_stream(uint4* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// 16 output bytes starting r bytes into the 32-byte little-endian pair (lo,hi).
// q = r>>2 selects the first 32-bit word, sh = 8*(r&3) the bit shift.
template <int Q>
__device__ __forceinline__ uint4 funnel16(const uint4& lo, const uint4& hi, uint32_t sh) {
  const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
  uint4 o;
  o.x = __funnelshift_r(w[Q + 0], w[Q + 1], sh);
  o.y = __funnelshift_r(w[Q + 1], w[Q + 2], sh);
  o.z = __funnelshift_r(w[Q + 2], w[Q + 3], sh);
  o.w = __funnelshift_r(w[Q + 3], w[Q + 4], sh);
  return o;
}

__device__ __forceinline__ uint4 shfl_down1(const uint4& v) {
  uint4 o;
  o.x = __shfl_down_sync(0xffffffffu, v.x, 1);
  o.y = __shfl_down_sync(0xffffffffu, v.y, 1);
  o.z = __shfl_down_sync(0xffffffffu, v.z, 1);
  o.w = __shfl_down_sync(0xffffffffu, v.w, 1);
  return o;
}

// Copy nvec 16-B vectors: dst is 16-B aligned, src = abase + r (abase aligned).
// Output vector i needs the aligned words W[i] and W[i+1].  Every lane loads
// its W[i] ONCE (coalesced 512 B per warp) and takes W[i+1] from its right
// neighbour with a shuffle; only lane 31 loads the extra halo word.  (Loading
// both words per lane re-fetched the shared sectors: ncu showed 18.7 GB of
// DRAM reads for 16.06 GB of payload, profiles/r01_shifted_path.md.)
template <int Q>
__device__ __forceinline__ void copy_shifted(const uint4* __restrict__ abase,
                                             uint4* __restrict__ dst, uint32_t nvec,
                                             uint32_t sh) {
  constexpr int T = kLsuThreads, U = kLsuUnroll;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t wbase = threadIdx.x - lane;  // first vector of this warp in the block row
  // W[nvec] (the halo of the last vector) shares a 16-B word with valid source
  // bytes, so indices <= nvec are readable.
  for (uint32_t base = wbase; base < nvec; base += U * T) {  // warp-uniform bounds
    uint4 lo[U], hi[U];
    // all global loads first (body + lane-31 halos), so one memory latency is
    // exposed per iteration, not two
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const uint32_t idx = base + j * T + lane;
      lo[j] = idx <= nvec ? ldg_stream(abase + idx) : make_uint4(0, 0, 0, 0);
      hi[j] = (lane == 31u && idx < nvec) ? ldg_cached(abase + idx + 1) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const uint4 nb = shfl_down1(lo[j]);
      if (lane != 31u) hi[j] = nb;
    }
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const uint32_t idx = base + j * T + lane;
      if (idx < nvec) stg_stream(dst + idx, funnel16<Q>(lo[j], hi[j], sh));
    }
  }
}

// CTA-wide copy of n bytes, any alignment on either side.
__device__ __forceinline__ void copy_range(const uint8_t* __restrict__ src,
                                           uint8_t* __restrict__ dst, uint32_t n) {
  constexpr int T = kLsuThreads, U = kLsuUnroll;
  // head: bring dst to 16-B alignment
  uint32_t head = (uint32_t)((16u - ((uintptr_t)dst & 15u)) & 15u);
  if (head > n) head = n;
  if (threadIdx.x < head) dst[threadIdx.x] = src[threadIdx.x];
  src += head;
  dst += head;
  n -= head;
  const uint32_t nvec = n >> 4;
  const uint32_t tail = n & 15u;
  if (threadIdx.x < tail) {
    const uint32_t o = (nvec << 4) + threadIdx.x;
    dst[o] = src[o];
  }
  if (nvec == 0) return;
  const uint32_t r = (uint32_t)((uintptr_t)src & 15u);
  uint4* __restrict__ d = reinterpret_cast<uint4*>(dst);
  if (r == 0) {
    const uint4* __restrict__ s = reinterpret_cast<const uint4*>(src);
    uint32_t i = threadIdx.x;
    for (; i + (U - 1) * T < nvec; i += U * T) {
      uint4 v[U];
#pragma unroll
      for (int j = 0; j < U; ++j) v[j] = ldg_stream(s + i + j * T);
#pragma unroll
      for (int j = 0; j < U; ++j) stg_stream(d + i + j * T, v[j]);
    }
    for (; i < nvec; i += T) stg_stream(d + i, ldg_stream(s + i));
  } else {
    // The two aligned words that straddle each output vector: the first starts
    // r bytes before src, the last ends (16-r) bytes after the body; both share
    // a 16-B word with at least one valid source byte, so no page is touched
    // that the source range does not already touch.
    const uint4* __restrict__ ab = reinterpret_cast<const uint4*>(src - r);
    const uint32_t sh = (r & 3u) * 8u;
    switch (r >> 2) {
      case 0: copy_shifted<0>(ab, d, nvec, sh); break;
      case 1: copy_shifted<1>(ab, d, nvec, sh); break;
      case 2: copy_shifted<2>(ab, d, nvec, sh); break;
      default: copy_shifted<3>(ab, d, nvec, sh); break;
    }
  }
}

// DIR 0: tensors -> arena (pack).  DIR 1: arena -> tensors (unpack).
template <int DIR>
__global__ void __launch_bounds__(kLsuThreads)
fc_copy_lsu(const FcItem* __restrict__ items, uint32_t n_items, uint8_t* __restrict__ arena) {
  for (uint32_t c = blockIdx.x; c < n_items; c += gridDim.x) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(items + c));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(items + c) + 1);
    uint8_t* t = reinterpret_cast<uint8_t*>(((uint64_t)a.y << 32) | a.x);
    uint8_t* ar = arena + (((uint64_t)a.w << 32) | a.z);
    const uint32_t n = b.x;
    if (DIR == 0)
      copy_range(t, ar, n);
    else
      copy_range(ar, t, n);
  }
}

// ---------------------------------------------------------------- TMA kernel --

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "FC_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra FC_DONE_%=;\n"
      "bra FC_WAIT_%=;\n"
      "FC_DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t smem_dst, const void* gsrc, uint32_t bytes,
                                         uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_dst),
      "l"(gsrc), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, uint32_t smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"(smem_src), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// Walks the tiles of the items owned by this CTA (items blockIdx.x, +gridDim.x, ...).
struct TileCursor {
  const FcItem* items;
  uint32_t n_items, item, off, n;
  uint64_t tptr, aoff;
  __device__ __forceinline__ void fetch() {
    if (item < n_items) {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(items + item));
      const uint4 b = __ldg(reinterpret_cast<const uint4*>(items + item) + 1);
      tptr = ((uint64_t)a.y << 32) | a.x;
      aoff = ((uint64_t)a.w << 32) | a.z;
      n = b.x;
    }
  }
  __device__ __forceinline__ void init(const FcItem* it, uint32_t cnt) {
    items = it;
    n_items = cnt;
    item = blockIdx.x;
    off = 0;
    fetch();
  }
  __device__ __forceinline__ bool valid() const { return item < n_items; }
  __device__ __forceinline__ uint32_t bytes(uint32_t tile) const {
    const uint32_t left = n - off;
    return left < tile ? left : tile;
  }
  __device__ __forceinline__ void advance(uint32_t tile) {
    off += tile;
    if (off >= n) {
      item += gridDim.x;
      off = 0;
      fetch();
    }
  }
};

// Every item handed to this kernel has tptr, aoff (and the arena base) 16-B
// aligned and nbytes a multiple of 16 — the plan builder guarantees it.
template <int DIR>
__global__ void __launch_bounds__(32)
fc_copy_tma(const FcItem* __restrict__ items, uint32_t n_items, uint8_t* __restrict__ arena,
            uint32_t tile, uint32_t stages) {
  extern __shared__ __align__(128) uint8_t fc_smem[];
  if (threadIdx.x != 0) return;
  const uint32_t smem_base = smem_u32(fc_smem);
  const uint32_t bar_base = smem_base + stages * tile;
  for (uint32_t s = 0; s < stages; ++s) mbar_init(bar_base + 8 * s, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

  TileCursor ld, st;
  ld.init(items, n_items);
  st.init(items, n_items);

  auto issue_load = [&](uint32_t stage) {
    const uint32_t nb = ld.bytes(tile);
    const void* g = DIR == 0 ? reinterpret_cast<const void*>(ld.tptr + ld.off)
                             : reinterpret_cast<const void*>(arena + ld.aoff + ld.off);
    mbar_expect_tx(bar_base + 8 * stage, nb);
    bulk_g2s(smem_base + stage * tile, g, nb, bar_base + 8 * stage);
    ld.advance(tile);
  };

  uint32_t primed = 0;
  while (primed < stages && ld.valid()) issue_load(primed++);

  for (uint32_t k = 0; st.valid(); ++k) {
    const uint32_t stage = k % stages;
    mbar_wait(bar_base + 8 * stage, (k / stages) & 1u);
    const uint32_t nb = st.bytes(tile);
    void* g = DIR == 0 ? reinterpret_cast<void*>(arena + st.aoff + st.off)
                       : reinterpret_cast<void*>(st.tptr + st.off);
    bulk_s2g(g, smem_base + stage * tile, nb);
    bulk_commit();
    st.advance(tile);
    if (k >= 1 && ld.valid()) {
      bulk_wait_read<1>();  // store k-1 (and older) no longer reads its stage
      issue_load((k - 1) % stages);
    }
  }
  bulk_wait_all();
}

// ---- TMA-fed byte-shift kernel ------------------------------------------------
// For ranges whose source and destination are NOT congruent mod 16 (everything
// behind a 4-byte optimizer `step` scalar in the reference's unpadded layout).
// Global traffic is all bulk-async and fully sector-efficient: the aligned
// source span of each 16 KiB destination tile (+ one 16-B halo word) is
// TMA-loaded into a shared-memory ring, 128 threads funnel-shift it from shared
// to shared (two LDS.128 + four SHF + one STS.128 per 16 B), and the aligned
// result is TMA-stored.  The LSU shifted path re-reads straddled sectors
// (18.7 GB of DRAM reads for 16.06 GB, 5.98 ms); this one does not.
constexpr int kShiftThreads = 256;
// shared memory: in_stages x (tile + 128 B halo slot) input ring, 2 x tile
// output double buffer, in_stages mbarriers
static inline size_t shift_smem_bytes(uint32_t tile, uint32_t in_stages) {
  return (size_t)in_stages * (tile + 128u) + 2u * (size_t)tile + 8u * in_stages;
}

struct ShiftCursor {
  const FcItem* items;
  uint32_t n_items, item, voff;  // voff: vectors of the body already consumed
  // per item (after the destination-aligning peel)
  const uint8_t* src;  // first body byte (r = src & 15)
  uint8_t* dst;        // 16-B aligned
  uint32_t nvec, r;
};

template <int DIR>
__device__ __forceinline__ void shift_fetch(ShiftCursor& c, uint8_t* arena) {
  while (c.item < c.n_items) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(c.items + c.item));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(c.items + c.item) + 1);
    uint8_t* t = reinterpret_cast<uint8_t*>(((uint64_t)a.y << 32) | a.x);
    uint8_t* ar = arena + (((uint64_t)a.w << 32) | a.z);
    const uint8_t* s = DIR == 0 ? t : ar;
    uint8_t* d = DIR == 0 ? ar : t;
    uint32_t n = b.x;
    uint32_t head = (uint32_t)((16u - ((uintptr_t)d & 15u)) & 15u);
    if (head > n) head = n;
    c.src = s + head;
    c.dst = d + head;
    c.nvec = (n - head) >> 4;
    c.r = (uint32_t)((uintptr_t)c.src & 15u);
    c.voff = 0;
    if (c.nvec) return;
    c.item += gridDim.x;  // nothing but peel bytes: the consumer side copies them
  }
}

// Head (< 16 B before the aligned body) and tail (< 16 B after it) of an item.
template <int DIR>
__device__ __forceinline__ void shift_peel(const FcItem* items, uint32_t item, uint8_t* arena) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(items + item));
  const uint4 b = __ldg(reinterpret_cast<const uint4*>(items + item) + 1);
  uint8_t* t = reinterpret_cast<uint8_t*>(((uint64_t)a.y << 32) | a.x);
  uint8_t* ar = arena + (((uint64_t)a.w << 32) | a.z);
  const uint8_t* s = DIR == 0 ? t : ar;
  uint8_t* d = DIR == 0 ? ar : t;
  const uint32_t n = b.x;
  uint32_t head = (uint32_t)((16u - ((uintptr_t)d & 15u)) & 15u);
  if (head > n) head = n;
  const uint32_t body = (n - head) & ~15u;
  const uint32_t tail = n - head - body;
  if (threadIdx.x < head) d[threadIdx.x] = s[threadIdx.x];
  if (threadIdx.x >= 16 && threadIdx.x - 16 < tail) {
    const uint32_t o = head + body + threadIdx.x - 16;
    d[o] = s[o];
  }
}

template <int Q>
__device__ __forceinline__ void shift_tile(const uint8_t* in, uint8_t* out, uint32_t nv,
                                           uint32_t sh) {
  const uint4* __restrict__ w = reinterpret_cast<const uint4*>(in);
  uint4* __restrict__ o = reinterpret_cast<uint4*>(out);
  for (uint32_t i = threadIdx.x; i < nv; i += kShiftThreads) o[i] = funnel16<Q>(w[i], w[i + 1], sh);
}

template <int DIR>
__global__ void __launch_bounds__(kShiftThreads)
fc_copy_tma_shift(const FcItem* __restrict__ items, uint32_t n_items,
                  uint8_t* __restrict__ arena, uint32_t kShiftTile, uint32_t kShiftStages) {
  extern __shared__ __align__(128) uint8_t fc_smem[];
  const uint32_t kShiftInStride = kShiftTile + 128u;
  uint8_t* in_base = fc_smem;
  uint8_t* out_base = fc_smem + kShiftStages * kShiftInStride;
  const uint32_t bar_base = smem_u32(out_base + 2u * kShiftTile);
  const uint32_t kTileVec = kShiftTile / 16;
  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < kShiftStages; ++s) mbar_init(bar_base + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  // Peel bytes of every item this CTA owns (tiny, done up front).
  for (uint32_t it = blockIdx.x; it < n_items; it += gridDim.x) shift_peel<DIR>(items, it, arena);

  // Load cursor lives in thread 0 only; the consume cursor is replicated.
  ShiftCursor ld, cs;
  cs.items = items; cs.n_items = n_items; cs.item = blockIdx.x;
  shift_fetch<DIR>(cs, arena);
  ld = cs;

  auto tile_vecs = [&](const ShiftCursor& c) {
    const uint32_t left = c.nvec - c.voff;
    return left < kTileVec ? left : kTileVec;
  };
  auto advance = [&](ShiftCursor& c) {
    c.voff += kTileVec;
    if (c.voff >= c.nvec) {
      c.item += gridDim.x;
      shift_fetch<DIR>(c, arena);
    }
  };
  auto issue_load = [&](uint32_t stage) {  // thread 0
    const uint32_t nv = tile_vecs(ld);
    const uint32_t nb = nv * 16u + (ld.r ? 16u : 0u);  // + halo word
    const void* g = ld.src - ld.r + (size_t)ld.voff * 16u;
    mbar_expect_tx(bar_base + 8 * stage, nb);
    bulk_g2s(smem_u32(in_base + stage * kShiftInStride), g, nb, bar_base + 8 * stage);
    advance(ld);
  };

  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < kShiftStages && ld.item < n_items; ++s) issue_load(s);
  }

  for (uint32_t k = 0; cs.item < n_items; ++k) {
    const uint32_t stage = k % kShiftStages;
    mbar_wait(bar_base + 8 * stage, (k / kShiftStages) & 1u);
    const uint32_t nv = tile_vecs(cs);
    const uint8_t* in = in_base + stage * kShiftInStride;
    uint8_t* out = out_base + (k & 1u) * kShiftTile;
    const uint32_t sh = (cs.r & 3u) * 8u;
    switch (cs.r >> 2) {
      case 0: shift_tile<0>(in, out, nv, sh); break;
      case 1: shift_tile<1>(in, out, nv, sh); break;
      case 2: shift_tile<2>(in, out, nv, sh); break;
      default: shift_tile<3>(in, out, nv, sh); break;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // STS -> bulk store
    __syncthreads();  // tile shifted; input stage fully read
    if (threadIdx.x == 0) {
      bulk_s2g(cs.dst + (size_t)cs.voff * 16u, smem_u32(out), nv * 16u);
      bulk_commit();
      if (ld.item < n_items) issue_load(stage);  // refill the stage just consumed
      bulk_wait_read<1>();  // store k-1 done reading: out[(k+1)&1] is free again
    }
    advance(cs);
    __syncthreads();
  }
  if (threadIdx.x == 0) bulk_wait_all();
}

// ----------------------------------------------------------------- host side --

struct fc_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t copy_stream = nullptr;
  uint8_t* arena = nullptr;
  uint64_t arena_bytes = 0;
  uint64_t arena_cap = 0;  // 0 = unlimited; else fc_arena_reserve never allocates more
  // tuning
  // tuning: defaults picked from the B200 sweep in profiles/r01_sweep.md
  int variant = FC_VARIANT_TMA;
  int lsu_ctas_per_sm = 2;
  int tma_ctas_per_sm = 1;
  int tma_stages = 2;
  int tma_tile = 96 << 10;
  int shift_ctas_per_sm = 2;
  int shift_tile = 16 << 10;
  int shift_stages = 2;
  // save pipeline state (one in flight)
  cudaEvent_t ev_pack_start = nullptr, ev_pack_end = nullptr, ev_drain_start = nullptr,
              ev_drain_end = nullptr;
  uint64_t ticket = 0;        // last issued
  bool save_inflight = false; // until observed complete
  // restore pipeline state
  cudaEvent_t ev_fill_start = nullptr, ev_fill_end = nullptr, ev_scatter_end = nullptr;
  bool restore_inflight = false;
  std::vector<void*> registered;
  // evidence counters: kernels launched / DMA copies enqueued by this context
  uint64_t n_kernels = 0, n_memcpys = 0;
  // drain pump (one thread per context, started lazily)
  std::thread pump;
  std::mutex mu;
  std::condition_variable cv;
  struct DrainJob {
    uint8_t* host = nullptr;
    std::vector<FcRun> runs;
    std::vector<FcSpan> spans;  // in-place save: DMA straight from the tensors
    bool direct = false;
    uint64_t arena_base = 0;  // arena byte 0 holds this segment offset (hybrid: the cut)
    uint64_t ticket = 0;
  };
  uint64_t direct_ticket = 0;  // last ticket with an in-place part (sources frozen until drained)
  uint64_t inplace_done_ticket = 0;  // last ticket whose in-place part has left the tensors
  uint64_t held_ticket = 0;  // the pump must not start the drain of this ticket yet
  std::deque<DrainJob> jobs;
  bool pump_stop = false;
  uint64_t drained_ticket = 0;  // last ticket whose bytes are all in host memory
  int drain_rc = FC_OK;         // sticky error of the pump
  std::string drain_err;
  cudaEvent_t ring[kDrainRing] = {};
  // device/pinned buffers whose cudaFree (a device-wide sync) is deferred to a
  // moment that synchronises anyway
  std::vector<void*> dead_dev, dead_pinned;
  uint64_t drain_piece = kDrainPiece;
  int drain_depth = kDrainDepth;
};

static void pump_main(fc_ctx* c) {
  cudaSetDevice(c->device);
  for (;;) {
    fc_ctx::DrainJob job;
    {
      std::unique_lock<std::mutex> lk(c->mu);
      c->cv.wait(lk, [&] {
        return c->pump_stop || (!c->jobs.empty() && c->jobs.front().ticket != c->held_ticket);
      });
      if (c->jobs.empty()) return;  // stop requested and nothing queued
      if (c->jobs.front().ticket == c->held_ticket) c->held_ticket = 0;  // stopping: drain anyway
      job = std::move(c->jobs.front());
      c->jobs.pop_front();
    }
    cudaError_t e = cudaStreamWaitEvent(c->copy_stream, c->ev_pack_end, 0);
    if (e == cudaSuccess) e = cudaEventRecord(c->ev_drain_start, c->copy_stream);
    const int depth = std::max(1, std::min(c->drain_depth, kDrainRing));
    uint64_t k = 0, copies = 0;
    if (job.direct) {
      // one batch (<= drain_piece bytes, <= 64 copies) in flight, then wait: same
      // pacing rule as below, small tensors share a batch
      uint64_t batch_bytes = 0;
      int batch_n = 0;
      for (const FcSpan& sp : job.spans) {
        for (uint64_t o = 0; o < sp.len && e == cudaSuccess; o += c->drain_piece) {
          const uint64_t len = std::min<uint64_t>(c->drain_piece, sp.len - o);
          e = cudaMemcpyAsync(job.host + sp.off + o, (const uint8_t*)(uintptr_t)sp.tptr + o, len,
                              cudaMemcpyDeviceToHost, c->copy_stream);
          ++copies;
          batch_bytes += len;
          if (e == cudaSuccess && (batch_bytes >= c->drain_piece || ++batch_n >= 64)) {
            e = cudaEventRecord(c->ring[0], c->copy_stream);
            if (e == cudaSuccess) e = cudaEventSynchronize(c->ring[0]);
            batch_bytes = 0;
            batch_n = 0;
          }
        }
      }
      if (e == cudaSuccess && (batch_bytes || batch_n)) {
        e = cudaEventRecord(c->ring[0], c->copy_stream);
        if (e == cudaSuccess) e = cudaEventSynchronize(c->ring[0]);
      }
      {
        // the tensors are free again (on error too: nothing reads them any more)
        std::lock_guard<std::mutex> lk(c->mu);
        c->inplace_done_ticket = job.ticket;
      }
      c->cv.notify_all();
    }
    for (const FcRun& r : job.runs) {
      for (uint64_t o = 0; o < r.len && e == cudaSuccess; o += c->drain_piece) {
        const uint64_t len = std::min<uint64_t>(c->drain_piece, r.len - o);
        if (k >= (uint64_t)depth) e = cudaEventSynchronize(c->ring[k % depth]);  // piece k-depth
        if (e == cudaSuccess)
          e = cudaMemcpyAsync(job.host + r.off + o, c->arena + (r.off - job.arena_base) + o, len,
                              cudaMemcpyDeviceToHost, c->copy_stream);
        if (e == cudaSuccess) e = cudaEventRecord(c->ring[k % depth], c->copy_stream);
        ++k;
        ++copies;
      }
    }
    if (e == cudaSuccess) e = cudaEventRecord(c->ev_drain_end, c->copy_stream);
    if (e == cudaSuccess) e = cudaEventSynchronize(c->ev_drain_end);
    {
      std::lock_guard<std::mutex> lk(c->mu);
      c->n_memcpys += copies;
      if (e != cudaSuccess) {
        c->drain_rc = FC_ECUDA;
        c->drain_err = std::string("drain pump: ") + cudaGetErrorString(e);
        (void)cudaGetLastError();
      }
      c->drained_ticket = job.ticket;
    }
    c->cv.notify_all();
  }
}

// One device-resident work table + its pinned staging copy (grow-only).
struct FcTable {
  FcItem* dev = nullptr;
  FcItem* pinned = nullptr;
  uint32_t cap = 0;
  uint32_t n = 0;
};

struct fc_plan {
  fc_ctx* ctx = nullptr;
  FcTable all;    // every byte, <= chunk pieces (LSU variant)
  FcTable bulk;   // 16-B congruent bodies (TMA variant)
  FcTable resid;  // heads and tails of congruent ranges (TMA variant)
  FcTable shift;  // ranges not congruent mod 16 (TMA variant: fc_copy_tma_shift)
  uint64_t payload = 0, arena_end = 0;
  uint32_t chunk = kDefaultChunk;
  std::vector<FcRun> runs;
  std::vector<FcItem> h_all;  // host copy of `all`, ascending arena offset (windowed mode)
  std::vector<FcSpan> spans;  // source ranges by arena offset (in-place save / restore)
  cudaEvent_t ev_upload = nullptr;  // last table upload
  cudaEvent_t ev_last_use = nullptr;  // last kernel that read the tables
};

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

#define FC_GUARD(ctx)                                                      \
  DeviceGuard _guard((ctx)->device);                                       \
  if (!_guard.ok) return fail(FC_ECUDA, "cudaSetDevice failed%s%s")

extern "C" int fc_ctx_create(int device, fc_ctx** out) {
  if (!out || device < 0) return fail(FC_EINVAL, "fc_ctx_create: bad argument%s%s");
  fc_ctx* c = new (std::nothrow) fc_ctx();
  if (!c) return fail(FC_ENOMEM, "fc_ctx_create: host alloc%s%s");
  c->device = device;
  DeviceGuard g(device);
  if (!g.ok) {
    delete c;
    return fail(FC_ECUDA, "cudaSetDevice failed%s%s");
  }
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    delete c;
    return fail(FC_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  }
  c->sm_count = prop.multiProcessorCount;
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  e = cudaStreamCreateWithPriority(&c->copy_stream, cudaStreamNonBlocking, hi);
  cudaEvent_t* evs[] = {&c->ev_pack_start, &c->ev_pack_end,  &c->ev_drain_start, &c->ev_drain_end,
                        &c->ev_fill_start, &c->ev_fill_end,  &c->ev_scatter_end};
  for (cudaEvent_t* ev : evs)
    if (e == cudaSuccess) e = cudaEventCreate(ev);
  // The pump spin-waits between pieces: wake-up latency matters with one piece
  // in flight (measured: 55.6 GB/s spinning vs 52.4 GB/s with a blocking wait
  // at 32 MiB pieces), and it costs one busy host core only while a checkpoint
  // drains (~0.3 s).  FC_DRAIN_SPIN=0 selects the blocking wait.
  const char* spin_env = getenv("FC_DRAIN_SPIN");
  const bool drain_spin = !(spin_env && spin_env[0] == '0');
  for (int i = 0; i < kDrainRing; ++i)
    if (e == cudaSuccess)
      e = cudaEventCreateWithFlags(
          &c->ring[i], cudaEventDisableTiming | (drain_spin ? 0 : cudaEventBlockingSync));
  if (e != cudaSuccess) {
    fc_ctx_destroy(c);
    return fail(FC_ECUDA, "fc_ctx_create: %s", cudaGetErrorString(e));
  }
  *out = c;
  return FC_OK;
}

extern "C" int fc_ctx_destroy(fc_ctx* c) {
  if (!c) return FC_OK;
  DeviceGuard g(c->device);
  if (c->pump.joinable()) {
    {
      std::lock_guard<std::mutex> lk(c->mu);
      c->pump_stop = true;
    }
    c->cv.notify_all();
    c->pump.join();
  }
  if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
  for (int i = 0; i < kDrainRing; ++i)
    if (c->ring[i]) cudaEventDestroy(c->ring[i]);
  for (void* p : c->registered) cudaHostUnregister(p);
  cudaEvent_t evs[] = {c->ev_pack_start, c->ev_pack_end, c->ev_drain_start, c->ev_drain_end,
                       c->ev_fill_start, c->ev_fill_end, c->ev_scatter_end};
  for (cudaEvent_t ev : evs)
    if (ev) cudaEventDestroy(ev);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->arena) cudaFree(c->arena);
  for (void* q : c->dead_dev) cudaFree(q);
  for (void* q : c->dead_pinned) cudaFreeHost(q);
  delete c;
  return FC_OK;
}

static int refresh_inflight(fc_ctx* c) {
  if (c->save_inflight) {
    std::lock_guard<std::mutex> lk(c->mu);
    if (c->drained_ticket >= c->ticket) c->save_inflight = false;
  }
  if (c->restore_inflight) {
    cudaError_t e = cudaEventQuery(c->ev_scatter_end);
    if (e == cudaSuccess)
      c->restore_inflight = false;
    else if (e != cudaErrorNotReady)
      return fail(FC_ECUDA, "cudaEventQuery(scatter): %s", cudaGetErrorString(e));
  }
  return FC_OK;
}

extern "C" int fc_arena_reserve(fc_ctx* c, uint64_t bytes) {
  if (!c) return fail(FC_EINVAL, "fc_arena_reserve: null ctx%s%s");
  FC_GUARD(c);
  if (c->arena_cap && bytes > c->arena_cap) bytes = c->arena_cap;
  if (bytes <= c->arena_bytes) return FC_OK;
  int rc = refresh_inflight(c);
  if (rc) return rc;
  if (c->save_inflight || c->restore_inflight)
    return fail(FC_EBUSY, "fc_arena_reserve: arena in use%s%s");
  if (c->arena) {
    FC_CUDA(cudaFree(c->arena));
    c->arena = nullptr;
    c->arena_bytes = 0;
  }
  for (void* q : c->dead_dev) cudaFree(q);  // cudaFree above synchronised already
  for (void* q : c->dead_pinned) cudaFreeHost(q);
  c->dead_dev.clear();
  c->dead_pinned.clear();
  // +32: the shifted load path may read one aligned 16-B word past the end.
  uint64_t want = ((bytes + 32 + 511) / 512) * 512;
  cudaError_t e = cudaMalloc(&c->arena, want);
  if (e != cudaSuccess) {
    c->arena = nullptr;
    (void)cudaGetLastError();
    return fail(FC_ENOMEM, "cudaMalloc(arena): %s", cudaGetErrorString(e));
  }
  c->arena_bytes = bytes;
  return FC_OK;
}

extern "C" int fc_set_arena_limit(fc_ctx* c, uint64_t bytes) {
  if (!c) return fail(FC_EINVAL, "fc_set_arena_limit: null ctx%s%s");
  if (bytes && bytes < (8ull << 20))
    return fail(FC_EINVAL, "fc_set_arena_limit: at least 8 MiB%s%s");
  c->arena_cap = bytes;
  return FC_OK;
}

extern "C" int fc_arena_info(fc_ctx* c, void** dev_ptr, uint64_t* bytes) {
  if (!c) return fail(FC_EINVAL, "fc_arena_info: null ctx%s%s");
  if (dev_ptr) *dev_ptr = c->arena;
  if (bytes) *bytes = c->arena_bytes;
  return FC_OK;
}

// NUMA node the GPU's PCIe root hangs off (-1 if unknown / single node).
static int gpu_numa_node(int device) {
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) {
    (void)cudaGetLastError();
    return -1;
  }
  for (char* q = bus; *q; ++q)
    if (*q >= 'A' && *q <= 'Z') *q = (char)(*q - 'A' + 'a');
  char path[128];
  snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
  FILE* f = fopen(path, "r");
  if (!f) return -1;
  int node = -1;
  if (fscanf(f, "%d", &node) != 1) node = -1;
  fclose(f);
  return node;
}

// Ask the kernel to back [host, host+bytes) with pages of `node` (the DMA
// target should be local to the GPU's socket: with 8 ranks draining at once
// the inter-socket link would otherwise carry half of the 440 GB/s).  Applies
// to pages not faulted in yet; best effort, failures are ignored.
static void prefer_numa_node(void* host, uint64_t bytes, int node) {
#if defined(SYS_mbind)
  if (node < 0 || node >= 1024) return;
  const long pg = sysconf(_SC_PAGESIZE);
  uintptr_t lo = (uintptr_t)host & ~((uintptr_t)pg - 1);
  uintptr_t hi = ((uintptr_t)host + bytes + pg - 1) & ~((uintptr_t)pg - 1);
  unsigned long mask[16] = {0};
  mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
  const int kMpolPreferred = 1;
  (void)syscall(SYS_mbind, (void*)lo, (unsigned long)(hi - lo), kMpolPreferred, mask,
                (unsigned long)(8 * sizeof(mask)), 0u);
#else
  (void)host; (void)bytes; (void)node;
#endif
}

struct PrefaultJob {
  uint8_t* p;
  size_t n;
};

static void* prefault_worker(void* arg) {
  PrefaultJob* j = static_cast<PrefaultJob*>(arg);
#ifdef MADV_POPULATE_WRITE
  if (madvise(j->p, j->n, MADV_POPULATE_WRITE) == 0) return nullptr;
#endif
  // Fallback: read-touch every page (never modifies an attached segment).
  const long pg = sysconf(_SC_PAGESIZE);
  volatile uint8_t sink = 0;
  for (size_t o = 0; o < j->n; o += (size_t)pg) sink ^= j->p[o];
  (void)sink;
  return nullptr;
}

extern "C" int fc_host_register(fc_ctx* c, void* host, uint64_t bytes, int prefault_threads) {
  if (!c || !host || bytes == 0) return fail(FC_EINVAL, "fc_host_register: bad argument%s%s");
  FC_GUARD(c);
  if (!getenv("FC_NO_NUMA")) prefer_numa_node(host, bytes, gpu_numa_node(c->device));
#ifdef MADV_HUGEPAGE
  // 2 MiB pages where the host allows them for this mapping (tmpfs: only with
  // transparent_hugepage/shmem_enabled = advise|always): fewer IOMMU / page-table
  // entries under the DMA.  Advice only; a refusal changes nothing.
  if (!getenv("FC_NO_HUGEPAGE")) {
    const uintptr_t pg = (uintptr_t)sysconf(_SC_PAGESIZE);
    uintptr_t lo = ((uintptr_t)host + pg - 1) & ~(pg - 1);
    uintptr_t hi = ((uintptr_t)host + bytes) & ~(pg - 1);
    if (hi > lo) (void)madvise((void*)lo, hi - lo, MADV_HUGEPAGE);
  }
#endif
  if (prefault_threads > 0) {
    const long pg = sysconf(_SC_PAGESIZE);
    int nt = std::min(prefault_threads, 64);
    uint64_t per = ((bytes / nt + pg - 1) / pg) * pg;
    if (per == 0) per = pg;
    std::vector<pthread_t> th;
    std::vector<PrefaultJob> jobs;
    jobs.reserve(nt);
    // mmap'd segments are page aligned; tolerate an unaligned start anyway
    uint8_t* base = static_cast<uint8_t*>(host);
    for (uint64_t o = 0; o < bytes; o += per) jobs.push_back({base + o, (size_t)std::min<uint64_t>(per, bytes - o)});
    th.resize(jobs.size());
    for (size_t i = 0; i < jobs.size(); ++i)
      if (pthread_create(&th[i], nullptr, prefault_worker, &jobs[i]) != 0) {
        prefault_worker(&jobs[i]);
        th[i] = 0;
      }
    for (size_t i = 0; i < jobs.size(); ++i)
      if (th[i]) pthread_join(th[i], nullptr);
  }
  cudaError_t e = cudaHostRegister(host, bytes, cudaHostRegisterPortable);
  if (e == cudaErrorHostMemoryAlreadyRegistered) {
    (void)cudaGetLastError();
    return FC_OK;
  }
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return fail(FC_ECUDA, "cudaHostRegister: %s", cudaGetErrorString(e));
  }
  c->registered.push_back(host);
  return FC_OK;
}

extern "C" int fc_host_unregister(fc_ctx* c, void* host) {
  if (!c || !host) return fail(FC_EINVAL, "fc_host_unregister: bad argument%s%s");
  FC_GUARD(c);
  auto it = std::find(c->registered.begin(), c->registered.end(), host);
  if (it == c->registered.end()) return FC_OK;
  // no DMA may still target the range: let the pump finish, then the stream
  {
    std::unique_lock<std::mutex> lk(c->mu);
    c->cv.wait(lk, [&] { return c->drained_ticket >= c->ticket; });
  }
  FC_CUDA(cudaStreamSynchronize(c->copy_stream));
  c->registered.erase(it);
  FC_CUDA(cudaHostUnregister(host));
  return FC_OK;
}

// Split [off, off+n) so interior boundaries are 128-B aligned in arena space.
static void split_range(std::vector<FcItem>& out, uint64_t tptr, uint64_t off, uint64_t n,
                        uint32_t chunk) {
  uint64_t first = chunk - (off & 127u);
  while (n > 0) {
    uint64_t len = std::min<uint64_t>(n, first);
    FcItem it;
    it.tptr = tptr;
    it.aoff = off;
    it.nbytes = (uint32_t)len;
    it.pad0 = 0;
    it.pad1 = 0;
    out.push_back(it);
    tptr += len;
    off += len;
    n -= len;
    first = chunk;
  }
}

// Put `v` into table `t` (device copy ordered on `s`).  sync=true: blocking
// cudaMemcpy from the vector (plan creation).  sync=false: staged through the
// table's pinned buffer and cudaMemcpyAsync on `s`, never a device-wide sync.
static int table_set(fc_ctx* c, FcTable& t, const std::vector<FcItem>& v, cudaStream_t s,
                     bool sync) {
  t.n = (uint32_t)v.size();
  if (v.empty()) return FC_OK;
  if (v.size() > t.cap) {
    uint32_t cap = (uint32_t)std::min<uint64_t>(0xFFFFFFF0ull, v.size() + v.size() / 4 + 16);
    FcItem* nd = nullptr;
    FcItem* np = nullptr;
    cudaError_t e = cudaMalloc(&nd, (size_t)cap * sizeof(FcItem));
    if (e == cudaSuccess) e = cudaHostAlloc(&np, (size_t)cap * sizeof(FcItem), cudaHostAllocDefault);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      if (nd) c->dead_dev.push_back(nd);
      return fail(FC_ENOMEM, "plan table alloc: %s", cudaGetErrorString(e));
    }
    if (t.dev) c->dead_dev.push_back(t.dev);  // may still be read by a queued kernel
    if (t.pinned) c->dead_pinned.push_back(t.pinned);
    t.dev = nd;
    t.pinned = np;
    t.cap = cap;
  }
  const size_t bytes = v.size() * sizeof(FcItem);
  if (sync) {
    FC_CUDA(cudaMemcpy(t.dev, v.data(), bytes, cudaMemcpyHostToDevice));
  } else {
    memcpy(t.pinned, v.data(), bytes);
    FC_CUDA(cudaMemcpyAsync(t.dev, t.pinned, bytes, cudaMemcpyHostToDevice, s));
  }
  return FC_OK;
}

// (Re)build the plan's runs and work tables from n ranges.
static int plan_fill(fc_plan* p, uint32_t n, const void* const* dev_ptrs,
                     const uint64_t* arena_off, const uint64_t* nbytes, cudaStream_t s,
                     bool sync) {
  fc_ctx* c = p->ctx;
  const uint32_t chunk_bytes = p->chunk;
  uint64_t payload = 0, arena_end = 0;
  std::vector<FcRun> ranges, runs;
  std::vector<FcSpan> spans;
  ranges.reserve(n);
  spans.reserve(n);
  for (uint32_t i = 0; i < n; ++i) {
    if (nbytes[i] == 0) continue;
    if (!dev_ptrs[i])
      return fail(FC_EINVAL, "plan: null device pointer for a non-empty tensor%s%s");
    if (arena_off[i] + nbytes[i] < arena_off[i]) return fail(FC_EINVAL, "plan: offset overflow%s%s");
    ranges.push_back({arena_off[i], nbytes[i]});
    spans.push_back({(uint64_t)(uintptr_t)dev_ptrs[i], arena_off[i], nbytes[i]});
    payload += nbytes[i];
    arena_end = std::max(arena_end, arena_off[i] + nbytes[i]);
  }
  std::sort(ranges.begin(), ranges.end(),
            [](const FcRun& a, const FcRun& b) { return a.off < b.off; });
  for (const FcRun& r : ranges) {
    if (!runs.empty()) {
      FcRun& last = runs.back();
      if (r.off < last.off + last.len)
        return fail(FC_EINVAL, "plan: tensors overlap in the arena%s%s");
      if (r.off == last.off + last.len) {
        last.len += r.len;
        continue;
      }
    }
    runs.push_back(r);
  }

  std::vector<FcItem> all, bulk, resid, shift;
  for (uint32_t i = 0; i < n; ++i) {
    uint64_t nb = nbytes[i];
    if (nb == 0) continue;
    uint64_t tp = (uint64_t)(uintptr_t)dev_ptrs[i];
    uint64_t off = arena_off[i];
    split_range(all, tp, off, nb, chunk_bytes);
    if (((tp - off) & 15u) != 0) {  // not congruent mod 16: byte-shift path
      // small ranges are not worth a TMA tile: leave them to the LSU kernel
      if (nb < 4096)
        split_range(resid, tp, off, nb, chunk_bytes);
      else
        split_range(shift, tp, off, nb, chunk_bytes);
      continue;
    }
    uint64_t head = std::min<uint64_t>((16u - (off & 15u)) & 15u, nb);
    if (head) split_range(resid, tp, off, head, chunk_bytes);
    uint64_t body = (nb - head) & ~15ull;
    if (body) split_range(bulk, tp + head, off + head, body, chunk_bytes);
    uint64_t tail = nb - head - body;
    if (tail) split_range(resid, tp + head + body, off + head + body, tail, chunk_bytes);
  }
  if (all.size() > 0xFFFFFFF0ull) return fail(FC_EINVAL, "plan: too many work items%s%s");
  std::stable_sort(all.begin(), all.end(),
                   [](const FcItem& a, const FcItem& b) { return a.aoff < b.aoff; });
  int rc = table_set(c, p->all, all, s, sync);
  if (!rc) rc = table_set(c, p->bulk, bulk, s, sync);
  if (!rc) rc = table_set(c, p->resid, resid, s, sync);
  if (!rc) rc = table_set(c, p->shift, shift, s, sync);
  if (rc) return rc;
  p->payload = payload;
  p->arena_end = arena_end;
  p->runs.swap(runs);
  p->h_all.swap(all);
  std::sort(spans.begin(), spans.end(),
            [](const FcSpan& a, const FcSpan& b) { return a.off < b.off; });
  // NOT merged even where tensor and arena addresses both continue: two tensors may
  // sit in adjacent but separate device allocations, and a cudaMemcpy must not
  // straddle allocations (kernels do not care, the copy API does).
  p->spans.swap(spans);
  return FC_OK;
}

extern "C" int fc_plan_create(fc_ctx* c, uint32_t n, const void* const* dev_ptrs,
                              const uint64_t* arena_off, const uint64_t* nbytes,
                              uint32_t chunk_bytes, fc_plan** out) {
  if (!c || !out || (n && (!dev_ptrs || !arena_off || !nbytes)))
    return fail(FC_EINVAL, "fc_plan_create: null argument%s%s");
  if (chunk_bytes == 0) chunk_bytes = kDefaultChunk;
  if (chunk_bytes < 4096 || (chunk_bytes & 127u) || chunk_bytes > (1u << 30))
    return fail(FC_EINVAL, "fc_plan_create: chunk_bytes must be a multiple of 128 in [4 KiB, 1 GiB]%s%s");
  FC_GUARD(c);
  fc_plan* p = new (std::nothrow) fc_plan();
  if (!p) return fail(FC_ENOMEM, "fc_plan_create: host alloc%s%s");
  p->ctx = c;
  p->chunk = chunk_bytes;
  cudaError_t e = cudaEventCreateWithFlags(&p->ev_upload, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_last_use, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    fc_plan_destroy(p);
    return fail(FC_ECUDA, "fc_plan_create: %s", cudaGetErrorString(e));
  }
  int rc = plan_fill(p, n, dev_ptrs, arena_off, nbytes, nullptr, /*sync=*/true);
  if (rc) {
    fc_plan_destroy(p);
    return rc;
  }
  *out = p;
  return FC_OK;
}

extern "C" int fc_plan_update(fc_plan* p, uint32_t n, const void* const* dev_ptrs,
                              const uint64_t* arena_off, const uint64_t* nbytes, void* stream) {
  if (!p || (n && (!dev_ptrs || !arena_off || !nbytes)))
    return fail(FC_EINVAL, "fc_plan_update: null argument%s%s");
  fc_ctx* c = p->ctx;
  FC_GUARD(c);
  int rc = refresh_inflight(c);
  if (rc) return rc;
  if (c->save_inflight || c->restore_inflight)
    return fail(FC_EBUSY, "fc_plan_update: a save/restore is still in flight%s%s");
  // the staging buffers are reused: the previous upload must have executed
  FC_CUDA(cudaEventSynchronize(p->ev_upload));
  cudaStream_t s = (cudaStream_t)stream;
  // a kernel of this plan queued on ANOTHER stream must not see the new table
  FC_CUDA(cudaStreamWaitEvent(s, p->ev_last_use, 0));
  rc = plan_fill(p, n, dev_ptrs, arena_off, nbytes, s, /*sync=*/false);
  if (rc) return rc;
  FC_CUDA(cudaEventRecord(p->ev_upload, s));
  return FC_OK;
}

extern "C" int fc_plan_destroy(fc_plan* p) {
  if (!p) return FC_OK;
  fc_ctx* c = p->ctx;
  DeviceGuard g(c->device);
  // Queued kernels may still read the tables: wait for the last one (not for
  // the whole device) and defer the cudaFree (which would sync everything).
  if (p->ev_last_use) cudaEventSynchronize(p->ev_last_use);
  if (p->ev_upload) cudaEventSynchronize(p->ev_upload);
  for (FcTable* t : {&p->all, &p->bulk, &p->resid, &p->shift}) {
    if (t->dev) c->dead_dev.push_back(t->dev);
    if (t->pinned) c->dead_pinned.push_back(t->pinned);
  }
  if (p->ev_upload) cudaEventDestroy(p->ev_upload);
  if (p->ev_last_use) cudaEventDestroy(p->ev_last_use);
  delete p;
  return FC_OK;
}

extern "C" int fc_plan_spans(const fc_plan* p, uint32_t* n_spans) {
  if (!p || !n_spans) return fail(FC_EINVAL, "fc_plan_spans: null argument%s%s");
  *n_spans = (uint32_t)p->spans.size();
  return FC_OK;
}

extern "C" int fc_plan_info(const fc_plan* p, uint64_t* payload_bytes, uint32_t* n_items,
                            uint32_t* n_runs, uint64_t* arena_end) {
  if (!p) return fail(FC_EINVAL, "fc_plan_info: null plan%s%s");
  if (payload_bytes) *payload_bytes = p->payload;
  if (n_items) *n_items = p->all.n;
  if (n_runs) *n_runs = (uint32_t)p->runs.size();
  if (arena_end) *arena_end = p->arena_end;
  return FC_OK;
}

extern "C" int fc_launch_count(fc_ctx* c, uint64_t* kernels, uint64_t* memcpys) {
  if (!c) return fail(FC_EINVAL, "fc_launch_count: null ctx%s%s");
  if (kernels) *kernels = c->n_kernels;
  if (memcpys) *memcpys = c->n_memcpys;
  return FC_OK;
}

extern "C" int fc_set_variant(fc_ctx* c, int variant) {
  if (!c || variant < FC_VARIANT_AUTO || variant > FC_VARIANT_TMA)
    return fail(FC_EINVAL, "fc_set_variant: bad argument%s%s");
  c->variant = variant == FC_VARIANT_AUTO ? FC_VARIANT_TMA : variant;
  return FC_OK;
}

extern "C" int fc_set_launch(fc_ctx* c, int lsu_ctas_per_sm, int tma_ctas_per_sm, int tma_stages,
                             int tma_tile_bytes) {
  if (!c) return fail(FC_EINVAL, "fc_set_launch: null ctx%s%s");
  if (lsu_ctas_per_sm < 0 || lsu_ctas_per_sm > 8 || tma_ctas_per_sm < 0 || tma_ctas_per_sm > 16 ||
      tma_stages < 0 || tma_stages == 1 || tma_stages > 32 || tma_tile_bytes < 0 ||
      (tma_tile_bytes & 15) || tma_tile_bytes > (128 << 10))
    return fail(FC_EINVAL, "fc_set_launch: out of range%s%s");
  int stages = tma_stages ? tma_stages : c->tma_stages;
  int tile = tma_tile_bytes ? tma_tile_bytes : c->tma_tile;
  if ((uint64_t)stages * tile + 8ull * stages > (227u << 10))
    return fail(FC_EINVAL, "fc_set_launch: ring exceeds 227 KB of shared memory%s%s");
  if (lsu_ctas_per_sm) c->lsu_ctas_per_sm = lsu_ctas_per_sm;
  if (tma_ctas_per_sm) c->tma_ctas_per_sm = tma_ctas_per_sm;
  c->tma_stages = stages;
  c->tma_tile = tile;
  return FC_OK;
}

template <int DIR>
static int launch_lsu(fc_ctx* c, const FcItem* items, uint32_t n, cudaStream_t s) {
  if (n == 0) return FC_OK;
  uint32_t grid = std::min<uint32_t>(n, (uint32_t)(c->sm_count * c->lsu_ctas_per_sm));
  fc_copy_lsu<DIR><<<grid, kLsuThreads, 0, s>>>(items, n, c->arena);
  FC_CUDA(cudaGetLastError());
  c->n_kernels += 1;
  return FC_OK;
}

template <int DIR>
static int launch_tma(fc_ctx* c, const FcItem* items, uint32_t n, cudaStream_t s) {
  if (n == 0) return FC_OK;
  const size_t smem = (size_t)c->tma_stages * c->tma_tile + 8u * c->tma_stages;
  FC_CUDA(cudaFuncSetAttribute(fc_copy_tma<DIR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  uint32_t grid = std::min<uint32_t>(n, (uint32_t)(c->sm_count * c->tma_ctas_per_sm));
  fc_copy_tma<DIR><<<grid, 32, smem, s>>>(items, n, c->arena, (uint32_t)c->tma_tile,
                                           (uint32_t)c->tma_stages);
  FC_CUDA(cudaGetLastError());
  c->n_kernels += 1;
  return FC_OK;
}

template <int DIR>
static int launch_shift(fc_ctx* c, const FcItem* items, uint32_t n, cudaStream_t s) {
  if (n == 0) return FC_OK;
  const size_t smem = shift_smem_bytes((uint32_t)c->shift_tile, (uint32_t)c->shift_stages);
  FC_CUDA(cudaFuncSetAttribute(fc_copy_tma_shift<DIR>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  uint32_t grid = std::min<uint32_t>(n, (uint32_t)(c->sm_count * c->shift_ctas_per_sm));
  fc_copy_tma_shift<DIR><<<grid, kShiftThreads, smem, s>>>(
      items, n, c->arena, (uint32_t)c->shift_tile, (uint32_t)c->shift_stages);
  FC_CUDA(cudaGetLastError());
  c->n_kernels += 1;
  return FC_OK;
}

template <int DIR>
static int launch_copy(fc_plan* p, cudaStream_t s, int variant) {
  fc_ctx* c = p->ctx;
  if (p->arena_end > c->arena_bytes)
    return fail(FC_EINVAL, "arena smaller than the plan: call fc_arena_reserve first%s%s");
  if (variant == FC_VARIANT_AUTO) variant = c->variant;
  // tables may have been (re)uploaded on another stream
  FC_CUDA(cudaStreamWaitEvent(s, p->ev_upload, 0));
  int rc;
  if (variant == FC_VARIANT_TMA) {
    rc = launch_tma<DIR>(c, p->bulk.dev, p->bulk.n, s);
    if (!rc) rc = launch_shift<DIR>(c, p->shift.dev, p->shift.n, s);
    if (!rc) rc = launch_lsu<DIR>(c, p->resid.dev, p->resid.n, s);
  } else {
    rc = launch_lsu<DIR>(c, p->all.dev, p->all.n, s);
  }
  if (rc) return rc;
  FC_CUDA(cudaEventRecord(p->ev_last_use, s));
  return FC_OK;
}

extern "C" int fc_pack_async(fc_plan* p, void* stream, int variant) {
  if (!p) return fail(FC_EINVAL, "fc_pack_async: null plan%s%s");
  FC_GUARD(p->ctx);
  return launch_copy<0>(p, (cudaStream_t)stream, variant);
}

extern "C" int fc_unpack_async(fc_plan* p, void* stream, int variant) {
  if (!p) return fail(FC_EINVAL, "fc_unpack_async: null plan%s%s");
  FC_GUARD(p->ctx);
  return launch_copy<1>(p, (cudaStream_t)stream, variant);
}

// ---- bounded arena ("windowed") mode ---------------------------------------------
// When the state does not fit a second time in HBM (arena smaller than the
// plan: fc_set_arena_limit, or the full-size cudaMalloc failed) the checkpoint is
// streamed through the arena window by window: gather the items of one window
// (LSU kernel over a slice of the offset-sorted table), DMA the window's bytes,
// next window.  The tensors must not change until the LAST window has been
// gathered, so this mode blocks the caller for the whole checkpoint — the
// reference's behaviour, at PCIe instead of pageable-copy speed.
struct Window {
  uint32_t i0, i1;      // item index range in h_all / d_all
  uint64_t base, end;   // arena byte range covered
};

static std::vector<Window> make_windows(const fc_plan* p, uint64_t arena_bytes) {
  std::vector<Window> w;
  const std::vector<FcItem>& it = p->h_all;
  uint32_t i = 0, n = (uint32_t)it.size();
  while (i < n) {
    Window cur{i, i, it[i].aoff, it[i].aoff};
    while (cur.i1 < n && it[cur.i1].aoff + it[cur.i1].nbytes - cur.base <= arena_bytes) {
      cur.end = std::max<uint64_t>(cur.end, it[cur.i1].aoff + it[cur.i1].nbytes);
      ++cur.i1;
    }
    if (cur.i1 == cur.i0) return {};  // one item larger than the arena: cannot happen (>= 8 MiB)
    w.push_back(cur);
    i = cur.i1;
  }
  return w;
}

// DMA the parts of the plan's runs that fall into [lo, hi), one piece at a time.
template <bool TO_HOST>
static int copy_window(fc_ctx* c, const fc_plan* p, uint8_t* host, uint64_t lo, uint64_t hi,
                       uint64_t base) {
  for (const FcRun& r : p->runs) {
    uint64_t a = std::max(lo, r.off), b = std::min(hi, r.off + r.len);
    for (uint64_t o = a; o < b; o += c->drain_piece) {
      const uint64_t len = std::min<uint64_t>(c->drain_piece, b - o);
      if (TO_HOST)
        FC_CUDA(cudaMemcpyAsync(host + o, c->arena + (o - base), len, cudaMemcpyDeviceToHost,
                                c->copy_stream));
      else
        FC_CUDA(cudaMemcpyAsync(c->arena + (o - base), host + o, len, cudaMemcpyHostToDevice,
                                c->copy_stream));
      c->n_memcpys += 1;
      FC_CUDA(cudaStreamSynchronize(c->copy_stream));  // paced: one piece in flight
    }
  }
  return FC_OK;
}

static int save_windowed(fc_plan* p, uint8_t* host, cudaStream_t cs) {
  fc_ctx* c = p->ctx;
  std::vector<Window> wins = make_windows(p, c->arena_bytes);
  if (wins.empty() && !p->h_all.empty())
    return fail(FC_EINVAL, "arena window too small for a work item%s%s");
  FC_CUDA(cudaStreamWaitEvent(cs, p->ev_upload, 0));
  FC_CUDA(cudaEventRecord(c->ev_pack_start, cs));
  bool first = true;
  for (const Window& w : wins) {
    // arena - base: the kernel adds the item's absolute arena offset
    uint32_t n = w.i1 - w.i0;
    uint32_t grid = std::min<uint32_t>(n, (uint32_t)(c->sm_count * c->lsu_ctas_per_sm));
    fc_copy_lsu<0><<<grid, kLsuThreads, 0, cs>>>(p->all.dev + w.i0, n, c->arena - w.base);
    FC_CUDA(cudaGetLastError());
    c->n_kernels += 1;
    FC_CUDA(cudaEventRecord(c->ev_pack_end, cs));
    FC_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_pack_end, 0));
    if (first) {
      FC_CUDA(cudaEventRecord(c->ev_drain_start, c->copy_stream));
      first = false;
    }
    int rc = copy_window<true>(c, p, host, w.base, w.end, w.base);
    if (rc) return rc;
    // the next gather overwrites the window: it must wait for this drain
    FC_CUDA(cudaEventRecord(c->ev_drain_end, c->copy_stream));
    FC_CUDA(cudaStreamWaitEvent(cs, c->ev_drain_end, 0));
  }
  if (first) FC_CUDA(cudaEventRecord(c->ev_drain_start, c->copy_stream));
  FC_CUDA(cudaEventRecord(c->ev_pack_end, cs));
  FC_CUDA(cudaEventRecord(p->ev_last_use, cs));
  FC_CUDA(cudaEventRecord(c->ev_drain_end, c->copy_stream));
  FC_CUDA(cudaEventSynchronize(c->ev_drain_end));
  FC_CUDA(cudaEventSynchronize(c->ev_pack_end));
  return FC_OK;
}

static int restore_windowed(fc_plan* p, const uint8_t* host, cudaStream_t s) {
  fc_ctx* c = p->ctx;
  std::vector<Window> wins = make_windows(p, c->arena_bytes);
  if (wins.empty() && !p->h_all.empty())
    return fail(FC_EINVAL, "arena window too small for a work item%s%s");
  FC_CUDA(cudaStreamWaitEvent(s, p->ev_upload, 0));
  FC_CUDA(cudaEventRecord(c->ev_fill_start, c->copy_stream));
  for (const Window& w : wins) {
    int rc = copy_window<false>(c, p, const_cast<uint8_t*>(host), w.base, w.end, w.base);
    if (rc) return rc;
    FC_CUDA(cudaEventRecord(c->ev_fill_end, c->copy_stream));
    FC_CUDA(cudaStreamWaitEvent(s, c->ev_fill_end, 0));
    uint32_t n = w.i1 - w.i0;
    uint32_t grid = std::min<uint32_t>(n, (uint32_t)(c->sm_count * c->lsu_ctas_per_sm));
    fc_copy_lsu<1><<<grid, kLsuThreads, 0, s>>>(p->all.dev + w.i0, n, c->arena - w.base);
    FC_CUDA(cudaGetLastError());
    c->n_kernels += 1;
    FC_CUDA(cudaEventRecord(c->ev_scatter_end, s));
    FC_CUDA(cudaEventSynchronize(c->ev_scatter_end));  // window is reused by the next fill
  }
  FC_CUDA(cudaEventRecord(c->ev_fill_end, c->copy_stream));
  FC_CUDA(cudaEventRecord(p->ev_last_use, s));
  FC_CUDA(cudaEventRecord(c->ev_scatter_end, s));
  return FC_OK;
}

static int save_async_impl(fc_plan* p, void* host_base, void* compute_stream, uint64_t* ticket,
                           bool hold);

extern "C" int fc_save_async(fc_plan* p, void* host_base, void* compute_stream, uint64_t* ticket) {
  return save_async_impl(p, host_base, compute_stream, ticket, false);
}

extern "C" int fc_save_async_held(fc_plan* p, void* host_base, void* compute_stream,
                                  uint64_t* ticket) {
  return save_async_impl(p, host_base, compute_stream, ticket, true);
}

extern "C" int fc_save_release(fc_ctx* c, uint64_t ticket) {
  if (!c) return fail(FC_EINVAL, "fc_save_release: null ctx%s%s");
  {
    std::lock_guard<std::mutex> lk(c->mu);
    if (c->held_ticket == ticket) c->held_ticket = 0;
  }
  c->cv.notify_all();
  return FC_OK;
}

static int save_async_impl(fc_plan* p, void* host_base, void* compute_stream, uint64_t* ticket,
                           bool hold) {
  if (!p || (!host_base && p->payload)) return fail(FC_EINVAL, "fc_save_async: null argument%s%s");
  fc_ctx* c = p->ctx;
  FC_GUARD(c);
  int rc = refresh_inflight(c);
  if (rc) return rc;
  if (c->save_inflight || c->restore_inflight)
    return fail(FC_EBUSY, "fc_save_async: previous save/restore still draining%s%s");
  cudaStream_t cs = (cudaStream_t)compute_stream;
  if (p->arena_end > c->arena_bytes) {
    if (c->arena_bytes < (8ull << 20))
      return fail(FC_EINVAL, "arena smaller than the plan: call fc_arena_reserve first%s%s");
    rc = save_windowed(p, static_cast<uint8_t*>(host_base), cs);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    c->ticket += 1;
    c->drained_ticket = c->ticket;  // already complete
    if (ticket) *ticket = c->ticket;
    return FC_OK;
  }
  FC_CUDA(cudaEventRecord(c->ev_pack_start, cs));
  rc = launch_copy<0>(p, cs, FC_VARIANT_AUTO);
  if (rc) return rc;
  FC_CUDA(cudaEventRecord(c->ev_pack_end, cs));
  // hand the drain to the pump thread (paced submission, see kDrainPiece)
  {
    std::lock_guard<std::mutex> lk(c->mu);
    if (c->drain_rc != FC_OK) return fail(c->drain_rc, "%s", c->drain_err.c_str());
    if (!c->pump.joinable()) c->pump = std::thread(pump_main, c);
    c->ticket += 1;
    fc_ctx::DrainJob job;
    job.host = static_cast<uint8_t*>(host_base);
    job.runs = p->runs;
    job.ticket = c->ticket;
    if (hold) c->held_ticket = c->ticket;
    c->jobs.push_back(std::move(job));
    c->save_inflight = true;
    if (ticket) *ticket = c->ticket;
  }
  c->cv.notify_all();
  return FC_OK;
}

// In-place part below `cut` (DMA from the tensors, drained first), snapshot part at
// and above it (LSU gather over a slice of the offset-sorted table into the arena,
// whose byte 0 stands for segment offset `cut`).
static int save_hybrid_impl(fc_plan* p, void* host_base, void* compute_stream, uint64_t cut,
                            int hold, uint64_t* ticket) {
  if (!p || (!host_base && p->payload))
    return fail(FC_EINVAL, "fc_save_hybrid_async: null argument%s%s");
  fc_ctx* c = p->ctx;
  FC_GUARD(c);
  int rc = refresh_inflight(c);
  if (rc) return rc;
  if (c->save_inflight || c->restore_inflight)
    return fail(FC_EBUSY, "fc_save_hybrid_async: previous save/restore still draining%s%s");
  cudaStream_t cs = (cudaStream_t)compute_stream;
  const std::vector<FcItem>& it = p->h_all;
  uint32_t n = (uint32_t)it.size();
  uint32_t i0 = (uint32_t)(std::lower_bound(it.begin(), it.end(), cut,
                                            [](const FcItem& a, uint64_t v) { return a.aoff < v; }) -
                           it.begin());
  if (i0 > 0 && it[i0 - 1].aoff + it[i0 - 1].nbytes > cut)
    return fail(FC_EINVAL, "fc_save_hybrid_async: cut is not a tensor boundary%s%s");
  if (i0 < n && p->arena_end - cut > c->arena_bytes)
    return fail(FC_EINVAL, "fc_save_hybrid_async: arena smaller than the snapshot part%s%s");
  // with no kernel the pair of events still orders the drain after everything
  // already queued on the training stream (the optimizer step that produced the values)
  FC_CUDA(cudaEventRecord(c->ev_pack_start, cs));
  if (i0 < n) {
    FC_CUDA(cudaStreamWaitEvent(cs, p->ev_upload, 0));
    uint32_t m = n - i0;
    uint32_t grid = std::min<uint32_t>(m, (uint32_t)(c->sm_count * c->lsu_ctas_per_sm));
    fc_copy_lsu<0><<<grid, kLsuThreads, 0, cs>>>(p->all.dev + i0, m, c->arena - cut);
    FC_CUDA(cudaGetLastError());
    c->n_kernels += 1;
    FC_CUDA(cudaEventRecord(p->ev_last_use, cs));
  }
  FC_CUDA(cudaEventRecord(c->ev_pack_end, cs));
  {
    std::lock_guard<std::mutex> lk(c->mu);
    if (c->drain_rc != FC_OK) return fail(c->drain_rc, "%s", c->drain_err.c_str());
    if (!c->pump.joinable()) c->pump = std::thread(pump_main, c);
    c->ticket += 1;
    fc_ctx::DrainJob job;
    job.host = static_cast<uint8_t*>(host_base);
    for (const FcSpan& sp : p->spans)
      if (sp.off < cut) job.spans.push_back(sp);
    for (const FcRun& r : p->runs) {
      uint64_t a = std::max(r.off, cut), b = r.off + r.len;
      if (b > a) job.runs.push_back({a, b - a});
    }
    job.direct = true;
    job.arena_base = cut;
    job.ticket = c->ticket;
    if (hold) c->held_ticket = c->ticket;
    c->direct_ticket = c->ticket;
    c->jobs.push_back(std::move(job));
    c->save_inflight = true;
    if (ticket) *ticket = c->ticket;
  }
  c->cv.notify_all();
  return FC_OK;
}

extern "C" int fc_save_hybrid_async(fc_plan* p, void* host_base, void* compute_stream,
                                    uint64_t cut, int hold, uint64_t* ticket) {
  return save_hybrid_impl(p, host_base, compute_stream, cut, hold, ticket);
}

extern "C" int fc_save_direct_async(fc_plan* p, void* host_base, void* compute_stream, int hold,
                                    uint64_t* ticket) {
  return save_hybrid_impl(p, host_base, compute_stream, ~0ull, hold, ticket);
}

extern "C" int fc_restore_direct_async(fc_plan* p, const void* host_base, void* stream) {
  if (!p || (!host_base && p->payload))
    return fail(FC_EINVAL, "fc_restore_direct_async: null argument%s%s");
  fc_ctx* c = p->ctx;
  FC_GUARD(c);
  int rc = refresh_inflight(c);
  if (rc) return rc;
  if (c->save_inflight || c->restore_inflight)
    return fail(FC_EBUSY, "fc_restore_direct_async: context busy%s%s");
  cudaStream_t s = (cudaStream_t)stream;
  const uint8_t* hb = static_cast<const uint8_t*>(host_base);
  // the targets may still be read by work queued on `stream`
  FC_CUDA(cudaEventRecord(c->ev_scatter_end, s));
  FC_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_scatter_end, 0));
  FC_CUDA(cudaEventRecord(c->ev_fill_start, c->copy_stream));
  for (const FcSpan& sp : p->spans)
    for (uint64_t o = 0; o < sp.len; o += kDmaPiece) {
      uint64_t len = std::min<uint64_t>(kDmaPiece, sp.len - o);
      FC_CUDA(cudaMemcpyAsync((uint8_t*)(uintptr_t)sp.tptr + o, hb + sp.off + o, len,
                              cudaMemcpyHostToDevice, c->copy_stream));
      c->n_memcpys += 1;
    }
  FC_CUDA(cudaEventRecord(c->ev_fill_end, c->copy_stream));
  FC_CUDA(cudaStreamWaitEvent(s, c->ev_fill_end, 0));
  FC_CUDA(cudaEventRecord(c->ev_scatter_end, s));
  c->restore_inflight = true;
  return FC_OK;
}

static int check_ticket(fc_ctx* c, uint64_t ticket, const char* who) {
  if (!c) return fail(FC_EINVAL, "%s: null ctx", who);
  if (ticket == 0 || ticket != c->ticket) return fail(FC_EINVAL, "%s: unknown ticket", who);
  return FC_OK;
}

static int drain_status(fc_ctx* c, uint64_t ticket, bool wait);

// FC_OK once nothing of save `ticket` reads the source tensors any more: its gather
// kernel (if any) has finished and its in-place part (if any) has been drained.
static int sources_status(fc_ctx* c, uint64_t ticket, bool wait) {
  {
    std::unique_lock<std::mutex> lk(c->mu);
    if (ticket == c->direct_ticket) {
      if (wait)
        c->cv.wait(lk, [&] { return c->inplace_done_ticket >= ticket || c->drain_rc != FC_OK; });
      if (c->drain_rc != FC_OK) return fail(c->drain_rc, "%s", c->drain_err.c_str());
      if (c->inplace_done_ticket < ticket) return FC_ENOTREADY;
    }
  }
  DeviceGuard g(c->device);
  if (!g.ok) return fail(FC_ECUDA, "cudaSetDevice failed%s%s");
  cudaError_t e = wait ? cudaEventSynchronize(c->ev_pack_end) : cudaEventQuery(c->ev_pack_end);
  if (e == cudaSuccess) return FC_OK;
  if (e == cudaErrorNotReady) return FC_ENOTREADY;
  return fail(FC_ECUDA, "cudaEvent(pack): %s", cudaGetErrorString(e));
}

extern "C" int fc_save_pack_done(fc_ctx* c, uint64_t ticket) {
  int rc = check_ticket(c, ticket, "fc_save_pack_done");
  if (rc) return rc;
  return sources_status(c, ticket, false);
}

extern "C" int fc_save_sources_wait(fc_ctx* c, uint64_t ticket) {
  int rc = check_ticket(c, ticket, "fc_save_sources_wait");
  if (rc) return rc;
  return sources_status(c, ticket, true);
}

static int drain_status(fc_ctx* c, uint64_t ticket, bool wait) {
  std::unique_lock<std::mutex> lk(c->mu);
  if (wait) c->cv.wait(lk, [&] { return c->drained_ticket >= ticket; });
  if (c->drained_ticket < ticket) return FC_ENOTREADY;
  if (c->drain_rc != FC_OK) return fail(c->drain_rc, "%s", c->drain_err.c_str());
  c->save_inflight = false;
  return FC_OK;
}

extern "C" int fc_save_poll(fc_ctx* c, uint64_t ticket) {
  int rc = check_ticket(c, ticket, "fc_save_poll");
  if (rc) return rc;
  return drain_status(c, ticket, false);
}

extern "C" int fc_save_wait(fc_ctx* c, uint64_t ticket) {
  int rc = check_ticket(c, ticket, "fc_save_wait");
  if (rc) return rc;
  return drain_status(c, ticket, true);
}

extern "C" int fc_save_timings(fc_ctx* c, uint64_t ticket, float* pack_ms, float* drain_ms,
                               float* total_ms) {
  int rc = check_ticket(c, ticket, "fc_save_timings");
  if (rc) return rc;
  rc = drain_status(c, ticket, true);
  if (rc) return rc;
  FC_GUARD(c);
  float t = 0.f;
  if (pack_ms) {
    FC_CUDA(cudaEventElapsedTime(&t, c->ev_pack_start, c->ev_pack_end));
    *pack_ms = t;
  }
  if (drain_ms) {
    FC_CUDA(cudaEventElapsedTime(&t, c->ev_drain_start, c->ev_drain_end));
    *drain_ms = t;
  }
  if (total_ms) {
    FC_CUDA(cudaEventElapsedTime(&t, c->ev_pack_start, c->ev_drain_end));
    *total_ms = t;
  }
  return FC_OK;
}

extern "C" int fc_set_shift_launch(fc_ctx* c, int ctas_per_sm, int in_stages, int tile_bytes) {
  if (!c) return fail(FC_EINVAL, "fc_set_shift_launch: null ctx%s%s");
  int cps = ctas_per_sm ? ctas_per_sm : c->shift_ctas_per_sm;
  int st = in_stages ? in_stages : c->shift_stages;
  int tile = tile_bytes ? tile_bytes : c->shift_tile;
  if (cps < 1 || cps > 8 || st < 2 || st > 8 || tile < 4096 || (tile & 127) ||
      shift_smem_bytes((uint32_t)tile, (uint32_t)st) > (227u << 10))
    return fail(FC_EINVAL, "fc_set_shift_launch: out of range%s%s");
  c->shift_ctas_per_sm = cps;
  c->shift_stages = st;
  c->shift_tile = tile;
  return FC_OK;
}

extern "C" int fc_set_drain(fc_ctx* c, uint64_t piece_bytes, int depth) {
  if (!c) return fail(FC_EINVAL, "fc_set_drain: null ctx%s%s");
  if ((piece_bytes && piece_bytes < (64u << 10)) || depth < 0 || depth > kDrainRing)
    return fail(FC_EINVAL, "fc_set_drain: piece >= 64 KiB, depth in [1, 8]%s%s");
  std::lock_guard<std::mutex> lk(c->mu);
  if (piece_bytes) c->drain_piece = piece_bytes;
  if (depth) c->drain_depth = depth;
  return FC_OK;
}

// ---- host-resident leaves ---------------------------------------------------
// CPU tensors inside a state_dict (optimizer step counters, RNG state, a whole
// CPU model in the gloo/CPU configuration) are already in host memory: they go
// straight into the segment with a (multi-threaded) memcpy, no device hop.

struct HostJob {
  uint8_t* dst;
  const void* const* src;
  const uint64_t* off;
  const uint64_t* nbytes;
  uint64_t first, last;
  uint64_t skip_first, trim_last;  // byte sub-range of the first / last range
};

static void* host_pack_worker(void* arg) {
  HostJob* j = static_cast<HostJob*>(arg);
  for (uint64_t i = j->first; i < j->last; ++i) {
    uint64_t lo = (i == j->first) ? j->skip_first : 0;
    uint64_t hi = (i + 1 == j->last) ? j->trim_last : j->nbytes[i];
    if (hi > lo)
      memcpy(j->dst + j->off[i] + lo, static_cast<const uint8_t*>(j->src[i]) + lo, hi - lo);
  }
  return nullptr;
}

extern "C" int fc_host_pack(void* dst_base, uint32_t n, const void* const* src,
                            const uint64_t* off, const uint64_t* nbytes, int threads) {
  if (!dst_base || (n && (!src || !off || !nbytes)))
    return fail(FC_EINVAL, "fc_host_pack: null argument%s%s");
  uint64_t total = 0;
  for (uint32_t i = 0; i < n; ++i) {
    if (nbytes[i] && !src[i]) return fail(FC_EINVAL, "fc_host_pack: null source%s%s");
    total += nbytes[i];
  }
  if (total == 0) return FC_OK;
  int nt = std::max(1, std::min(threads, 64));
  if (total < (8ull << 20)) nt = 1;  // not worth a thread
  // equal BYTE shares: a worker may start/stop in the middle of a range
  std::vector<HostJob> jobs;
  uint64_t share = (total + nt - 1) / nt;
  uint32_t i = 0;
  uint64_t inner = 0;  // bytes of range i already assigned
  while (i < n) {
    HostJob j{static_cast<uint8_t*>(dst_base), src, off, nbytes, i, i, inner, 0};
    uint64_t need = share;
    while (i < n && need > 0) {
      uint64_t left = nbytes[i] - inner;
      if (left <= need) {
        need -= left;
        j.trim_last = nbytes[i];
        ++i;
        inner = 0;
      } else {
        inner += need;
        j.trim_last = inner;
        need = 0;
        j.last = i + 1;
        break;
      }
      j.last = i;
    }
    if (j.last > j.first) jobs.push_back(j);
  }
  std::vector<pthread_t> th(jobs.size());
  for (size_t k = 0; k < jobs.size(); ++k) {
    if (k + 1 == jobs.size() || pthread_create(&th[k], nullptr, host_pack_worker, &jobs[k]) != 0) {
      host_pack_worker(&jobs[k]);
      th[k] = 0;
    }
  }
  for (size_t k = 0; k < jobs.size(); ++k)
    if (th[k]) pthread_join(th[k], nullptr);
  return FC_OK;
}

extern "C" int fc_restore_async(fc_plan* p, const void* host_base, void* stream) {
  if (!p || (!host_base && p->payload)) return fail(FC_EINVAL, "fc_restore_async: null argument%s%s");
  fc_ctx* c = p->ctx;
  FC_GUARD(c);
  int rc = refresh_inflight(c);
  if (rc) return rc;
  if (c->save_inflight || c->restore_inflight)
    return fail(FC_EBUSY, "fc_restore_async: arena busy%s%s");
  if (p->arena_end > c->arena_bytes) {
    if (c->arena_bytes < (8ull << 20))
      return fail(FC_EINVAL, "arena smaller than the plan: call fc_arena_reserve first%s%s");
    rc = restore_windowed(p, static_cast<const uint8_t*>(host_base), (cudaStream_t)stream);
    if (rc) return rc;
    c->restore_inflight = true;
    return FC_OK;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const uint8_t* hb = static_cast<const uint8_t*>(host_base);
  FC_CUDA(cudaEventRecord(c->ev_fill_start, c->copy_stream));
  for (const FcRun& r : p->runs)
    for (uint64_t o = 0; o < r.len; o += kDmaPiece) {
      uint64_t len = std::min<uint64_t>(kDmaPiece, r.len - o);
      FC_CUDA(cudaMemcpyAsync(c->arena + r.off + o, hb + r.off + o, len, cudaMemcpyHostToDevice,
                              c->copy_stream));
      c->n_memcpys += 1;
    }
  FC_CUDA(cudaEventRecord(c->ev_fill_end, c->copy_stream));
  FC_CUDA(cudaStreamWaitEvent(s, c->ev_fill_end, 0));
  rc = launch_copy<1>(p, s, FC_VARIANT_AUTO);
  if (rc) return rc;
  FC_CUDA(cudaEventRecord(c->ev_scatter_end, s));
  c->restore_inflight = true;
  return FC_OK;
}

extern "C" int fc_restore_wait(fc_ctx* c) {
  if (!c) return fail(FC_EINVAL, "fc_restore_wait: null ctx%s%s");
  FC_GUARD(c);
  FC_CUDA(cudaEventSynchronize(c->ev_scatter_end));
  c->restore_inflight = false;
  return FC_OK;
}

extern "C" int fc_restore_timings(fc_ctx* c, float* fill_ms, float* scatter_ms, float* total_ms) {
  if (!c) return fail(FC_EINVAL, "fc_restore_timings: null ctx%s%s");
  FC_GUARD(c);
  FC_CUDA(cudaEventSynchronize(c->ev_scatter_end));
  c->restore_inflight = false;
  float t = 0.f;
  if (fill_ms) {
    FC_CUDA(cudaEventElapsedTime(&t, c->ev_fill_start, c->ev_fill_end));
    *fill_ms = t;
  }
  if (scatter_ms) {
    FC_CUDA(cudaEventElapsedTime(&t, c->ev_fill_end, c->ev_scatter_end));
    *scatter_ms = t;
  }
  if (total_ms) {
    FC_CUDA(cudaEventElapsedTime(&t, c->ev_fill_start, c->ev_scatter_end));
    *total_ms = t;
  }
  return FC_OK;
}
