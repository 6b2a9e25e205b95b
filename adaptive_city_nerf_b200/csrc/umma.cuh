// Thin inline-PTX layer over Blackwell's 5th-gen tensor cores (tcgen05 / TMEM / mbarrier),
// sm_100a only.  Layout conventions used by every kernel that includes this file:
//
//   Operand tiles live in shared memory in the canonical NO-SWIZZLE K-major UMMA layout
//   (cute::UMMA::LayoutType::SWIZZLE_NONE, "INTERLEAVE"): the tile is cut into core matrices of
//   8 rows x 16 bytes (8 fp16), each stored as 128 contiguous bytes.  With KC = K/8 chunks,
//       byte offset of element (row r, col k) = (r>>3)*SBO + (k>>3)*LBO + (r&7)*16 + (k&7)*2,
//       LBO = 128 (next K chunk), SBO = KC*128 (next group of 8 rows).
//   A thread that owns row r writes its 16-byte chunks at stride 128 B; the 8 threads of a
//   quarter-warp (same r>>3) cover 128 contiguous bytes, so the stores are bank-conflict free.
//   The same bytes viewed as an MN-major operand (rows become K) have LBO and SBO swapped.
//
//   Accumulators (M = 128) live in TMEM: lane = row, column = n.  Warp w of a 128-thread CTA
//   reads lanes 32w..32w+31 with tcgen05.ld.32x32b, i.e. thread t reads row t.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- shared-memory (matrix) descriptor, cute::UMMA::SmemDescriptor bit layout -----------------
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);          // [0,14)  start address >> 4
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;     // [16,30) leading-dim byte offset >> 4
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;     // [32,46) stride-dim byte offset >> 4
    d |= (uint64_t)1 << 46;                                // [46,48) descriptor version = 1 (sm_100)
    return d;                                              // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE
}

// ---- instruction descriptor, cute::UMMA::InstrDescriptor (kind::f16, fp16 x fp16 -> fp32) -----
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, bool a_mn_major = false, bool b_mn_major = false) {
    return (1u << 4)                                  // c_format = F32
         | (0u << 7) | (0u << 10)                     // a_format = b_format = F16
         | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16)
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (M = 128 rows = lanes, K = 16 -> 8 columns of packed 16-bit pairs)
// is read from tensor memory.  Issued by ONE thread.
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}

// Arrive on an mbarrier once every previously issued MMA of this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// ---- TMEM management (warp-collective) ------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_in_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(dst_in_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (tensor core operand fetch)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// TMEM -> registers: this warp's 32 lanes x 16 consecutive 32-bit columns (thread = lane = row).
__device__ __forceinline__ void ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- mbarrier -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// Bounded wait (wall time, see mbar_wait_a below): a descriptor bug must surface as a trap (CUDA error), never as a hung GPU.
__device__ __forceinline__ void mbar_wait_a(uint32_t addr, uint32_t parity);
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { mbar_wait_a(smem_u32(bar), parity); }

// ---- address-based variants (32-bit shared-space addresses keep every access STS/LDS and every MMA operand in
// uniform registers; generic pointers into dynamic shared memory made ptxas emit generic LD/ST) -------------
__device__ __forceinline__ void mbar_init_a(uint32_t addr, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void commit_a(uint32_t addr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(addr) : "memory");
}
// Bounded wait: a protocol bug must surface as a trap (CUDA error), never as a hung GPU.  The bound is WALL TIME (4 s of
// %globaltimer, checked every 2^14 polls), not a poll count: how long one mbarrier.try_wait takes before it reports "not
// yet" is implementation-defined, and a poll-count bound that is safe for one code shape is not for another -- with the loop
// below kept rolled (it must be: unrolled four times at ~40 wait sites it was half of the fused backward's code, and the two
// roles of the warp-specialised kernels run different code at the same time: 8.51 -> 8.18 ms from the pragma alone) 2^22
// polls elapsed in legitimate waits about once in 10^2 training steps queued back to back.
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mbar_wait_a(uint32_t addr, uint32_t parity) {
    unsigned long long t0 = 0;
#pragma unroll 1
    for (uint32_t spin = 0;; ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
        if ((spin & 0x3FFFu) == 0x3FFFu) {
            const unsigned long long now = global_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ull) __trap();
        }
    }
}
// true in exactly one lane of a converged warp; the form ptxas recognises as "single thread" for tcgen05 issue
__device__ __forceinline__ bool elect_one() {
    uint32_t p;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(p));
    return p != 0;
}
// named barrier over `nthreads` threads (id 1..15; 0 is __syncthreads)
__device__ __forceinline__ void bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" :: "r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
// registers -> TMEM: 16 consecutive columns of this thread's lane
__device__ __forceinline__ void st16_zero(uint32_t taddr) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};"
        :: "r"(taddr), "r"(0u) : "memory");
}
// registers -> TMEM: 8 / 16 consecutive 32-bit columns of this thread's lane
__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- canonical-layout stores ----------------------------------------------------------------
// byte offset of the 16-byte chunk (row r, K chunk c) in a tile whose row-group stride is sbo
__device__ __forceinline__ uint32_t chunk_off(int r, int c, uint32_t sbo) {
    return (uint32_t)(r >> 3) * sbo + (uint32_t)c * 128u + (uint32_t)(r & 7) * 16u;
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ void st_chunk(uint8_t* tile, int r, int c, uint32_t sbo, const float* v8) {
    uint4 q = make_uint4(pack_h2(v8[0], v8[1]), pack_h2(v8[2], v8[3]), pack_h2(v8[4], v8[5]), pack_h2(v8[6], v8[7]));
    *reinterpret_cast<uint4*>(tile + chunk_off(r, c, sbo)) = q;
}

}  // namespace umma
