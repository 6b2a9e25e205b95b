// Fully fused per-expert forward (SURVEY 8b acn_render_expert_fwd): world->unit, multiresolution hash encode and the
// field MLPs in ONE persistent kernel -- nerfs/ray_rendering.py:317-325 -> models/inr/meta_ngp.py:226-241 ->
// models/encodings.py:293-381.  The (P, L*F) fp16 encoding never makes the HBM round trip between an encode kernel and
// an MLP kernel: it is written (optionally, for the backward) but not read back.
//
// One CTA per SM, warp-specialised:
//   * XNP producer warps run the hash encode of hashgrid.cu (thread = point, 8 gathers per level, the same fused-lerp
//     arithmetic as k_hashgrid_fwd<2, __half>: bit-identical rows) and store each point's fp16 row straight into a
//     shared-memory tile in the canonical K-major UMMA layout -- i.e. as the A operand of the first trunk layer;
//   * XNC consumer warpgroups run the MLP chain of k_field_fwd_mma: layer 1 takes A from that shared-memory tile
//     (tcgen05.mma, both operands in shared memory), every later layer takes A from TENSOR MEMORY where the previous
//     epilogue left it (field_mma.cu, "forward").
//   A ring of XST tiles decouples them: full[s] (4 producer warps arrive after fence.proxy.async) and empty[s] (arrived by
//   the tcgen05.commit that follows the layer-1 MMAs, i.e. when the tensor core has finished reading the tile).
//
// Optional staging of the coarsest levels (north star, stage 2: "stages per-level table tiles through shared memory"):
// in this reference EVERY level is hashed (models/encodings.py:308-316), so a level has no contiguous tile in the
// table; the CTA instead gathers the (res+1)^3 rows of the levels that fit (level 0: 17^3 rows = 39 KB, level 1: 23^3
// rows = 97 KB at the BASELINE grid) ONCE into a dense x-major lattice in shared memory, and the producers read those
// levels with LDS at lattice addresses (no hash, no L1/L2 traffic).  Values are copies, so the rows stay bit-identical.
//
// Row order.  Sample-major (explicit points, buckets, shuffled training rays): tile g holds points [128 g, 128 g + 128).
// Ray-major (frames, acn_rays_coherent): tile g = (ray block, sample group) holds 4 consecutive samples of 32 adjacent
// rays, one sample per producer warp, so the lanes of a warp share cells up to the mid levels (hashgrid.cu).
#include "field_mma.cuh"

namespace {

constexpr int XNC = 2;                       // consumer warpgroups
constexpr uint32_t XTMEM_COLS = XNC * 128;
// NP producer warps: NP / 4 tiles in production at a time, one tile stage per production slot (the slots stagger
// themselves after the first round).  16 producer + 8 consumer warps = 768 threads leave every thread 80 registers.
template <int NP> struct XCfg {
    static constexpr int threads = XNC * 128 + NP * 32;
    static constexpr int slots = NP / 4;
    // a stage's tiles must all go to the SAME consumer warpgroup (an mbarrier waiter may be at most one phase behind):
    // the stage count is kept a multiple of the XNC consumer warpgroups
    static constexpr int stages = (NP / 4) % XNC == 0 ? NP / 4 : XNC * (NP / 4);
};
constexpr int XNP = 16;

template <int E, int XST> struct XMap {
    static constexpr uint32_t w = 0;
    static constexpr uint32_t stages = (wmap(E).end + 1023u) & ~1023u;
    static constexpr uint32_t tile_bytes = TM * E * 2;
    static constexpr uint32_t bars = stages + XST * tile_bytes;          // full[XST], empty[XST], done[XNC]
    static constexpr uint32_t tmem_ptr = bars + (2 * XST + XNC) * 8u;
    static constexpr uint32_t res = tmem_ptr + 16u;                      // float resolution per level (16)
    static constexpr uint32_t lattice = (res + 64u + 127u) & ~127u;      // staged coarse levels (float2 per lattice node)
};

struct XArgs {
    const float* x; int xs;                           // explicit positions (P,>=3), or
    const float* rays; const float* t; int S;         // rays (N,8) + t (N,S)
    int64_t N;
    int ray_major; const int32_t* ray_major_dev;
    const float* box6;
    const float* table; int L; int log2T; const int32_t* res; int interp;
    int staged_levels; uint32_t lattice_off[2];       // levels [0, staged_levels) live in shared memory at these byte offsets
};

__device__ __forceinline__ void mbar_arrive(uint32_t addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(addr) : "memory");
}
__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
    return v;
}

// one level, F = 2, fp16 flavour of hashgrid_encode_point (fused lerps; same expressions, same operand order)
__device__ __forceinline__ float2 blend8(const float2* f, const GridCell& g) {
    float2 o;
    {
        const float c00 = fmaf(g.wx, f[4].x - f[0].x, f[0].x), c01 = fmaf(g.wx, f[5].x - f[1].x, f[1].x);
        const float c10 = fmaf(g.wx, f[6].x - f[2].x, f[2].x), c11 = fmaf(g.wx, f[7].x - f[3].x, f[3].x);
        const float c0 = fmaf(g.wy, c10 - c00, c00), c1 = fmaf(g.wy, c11 - c01, c01);
        o.x = fmaf(g.wz, c1 - c0, c0);
    }
    {
        const float c00 = fmaf(g.wx, f[4].y - f[0].y, f[0].y), c01 = fmaf(g.wx, f[5].y - f[1].y, f[1].y);
        const float c10 = fmaf(g.wx, f[6].y - f[2].y, f[2].y), c11 = fmaf(g.wx, f[7].y - f[3].y, f[3].y);
        const float c0 = fmaf(g.wy, c10 - c00, c00), c1 = fmaf(g.wy, c11 - c01, c01);
        o.y = fmaf(g.wz, c1 - c0, c0);
    }
    return o;
}

// One level from the table: eight 8-byte gathers; the x-neighbours of an even x0 are rows r and r^1 (the hash is
// x ^ y*p1 ^ z*p2), so when the WHOLE warp sits on even x0 (frames; hashgrid.cu) four aligned 16-byte loads fetch them.
// Measured and dropped (profiles/r02_v5_fused_fwd_variants.txt): per-lane pairs with the second load predicated on an odd
// x0 (6 instead of 8 requests per lane and level, but 3.91 vs 3.76 ms: four more registers per load in flight), and 24
// producer warps at 48 registers with setmaxnreg (4.19 ms) -- the gathers are bound in the L1 data pipe, not by latency.
__device__ __forceinline__ float2 level_from_table(const float* __restrict__ lt, const GridCell& g, uint32_t mask) {
    float2 f[8];
    if (__all_sync(0xffffffffu, !(g.x0 & 1u))) {
        const float4* lt4 = reinterpret_cast<const float4*>(lt);
        const uint32_t yp0 = g.y0 * 2654435761u, yp1 = yp0 + 2654435761u;
        const uint32_t zp0 = g.z0 * 805459861u, zp1 = zp0 + 805459861u;
#pragma unroll
        for (int yz = 0; yz < 4; ++yz) {
            const uint32_t h = ((yz & 2) ? yp1 : yp0) ^ ((yz & 1) ? zp1 : zp0);
            const uint32_t r0 = (g.x0 ^ h) & mask;
            const float4 v = __ldg(lt4 + (r0 >> 1));
            const bool odd = r0 & 1u;
            f[yz] = odd ? make_float2(v.z, v.w) : make_float2(v.x, v.y);
            f[4 + yz] = odd ? make_float2(v.x, v.y) : make_float2(v.z, v.w);
        }
    } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) f[c] = __ldg(reinterpret_cast<const float2*>(lt) + grid_corner_row(g, c, mask));
    }
    return blend8(f, g);
}

// staged level: dense lattice of (R+1)^3 nodes, node (x,y,z) at ((x*(R+1) + y)*(R+1) + z) * 8 bytes
__device__ __forceinline__ float2 level_from_lattice(uint32_t base, int R1, const GridCell& g) {
    float2 f[8];
    // unit positions are clamped to [1e-6, 1 - 1e-6], so x0 <= R - 1 and every corner is a lattice node; the min()
    // only guards NaN positions (which produce NaN weights, hence NaN features, as the table path does)
    const uint32_t x0 = min(g.x0, (uint32_t)(R1 - 2)), y0 = min(g.y0, (uint32_t)(R1 - 2)), z0 = min(g.z0, (uint32_t)(R1 - 2));
    const uint32_t a000 = base + (((x0 * R1) + y0) * R1 + z0) * 8u;
    const uint32_t dy = (uint32_t)R1 * 8u, dx = dy * (uint32_t)R1;
#pragma unroll
    for (int c = 0; c < 8; ++c) f[c] = lds_f2(a000 + ((c >> 2) & 1) * dx + ((c >> 1) & 1) * dy + (c & 1) * 8u);
    return blend8(f, g);
}

template <int E, int NP>
__global__ void __launch_bounds__(XCfg<NP>::threads, 1) k_expert_fwd(
    XArgs a, const float* __restrict__ dirs, int dstride, int dgroup, int64_t P, int G, acn_field_weights w,
    __half* __restrict__ enc_out, float4* __restrict__ rgb_sigma, const int32_t* __restrict__ range)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    if (range) {            // rows [range[0], range[1]) only (an expert's bucket): P was just the launch's upper bound
        const int64_t r0 = __ldg(range);
        P = __ldg(range + 1) - r0;
        a.x += r0 * a.xs; dirs += r0 * dstride; rgb_sigma += r0;
        if (enc_out) enc_out += r0 * E;
    }
    if (P <= 0) return;     // an empty bucket (the host never reads the counts): leave before staging weights / allocating TMEM
    using CF = XCfg<NP>;
    constexpr int XST = CF::stages, XSLOTS = CF::slots, XTHREADS = CF::threads;
    using M = XMap<E, XST>;
    constexpr WMap wm = wmap(E);
    constexpr int EC = E / 8;
    const uint32_t sb = umma::smem_u32(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t full0 = sb + M::bars, empty0 = full0 + 8u * XST, done0 = empty0 + 8u * XST;

    stage_weights<E>(w, G, sb + M::w);
    if (tid == 0) {
        for (int i = 0; i < XST; ++i) { umma::mbar_init_a(full0 + 8 * i, 4); umma::mbar_init_a(empty0 + 8 * i, 1); }
        for (int i = 0; i < XNC; ++i) umma::mbar_init_a(done0 + 8 * i, 1);
        umma::fence_mbar_init();
    }
    if (tid < 16) umma::sts_f32(sb + M::res + 4u * tid, tid < a.L ? (float)__ldg(a.res + tid) : 1.0f);
    const uint32_t hmask = (1u << a.log2T) - 1u;
    // stage the coarse levels that fit: one hashed gather per lattice node, once per CTA
    for (int l = 0; l < a.staged_levels; ++l) {
        const int R1 = __ldg(a.res + l) + 1;
        const float2* lt = reinterpret_cast<const float2*>(a.table) + ((size_t)l << a.log2T);
        const uint32_t base = sb + M::lattice + a.lattice_off[l];
        for (int i = tid; i < R1 * R1 * R1; i += XTHREADS) {
            const int z = i % R1, y = (i / R1) % R1, x = i / (R1 * R1);
            const float2 v = __ldg(lt + grid_hash((uint32_t)x, (uint32_t)y, (uint32_t)z, hmask));
            asm volatile("st.shared.v2.f32 [%0], {%1,%2};" :: "r"(base + 8u * (uint32_t)i), "f"(v.x), "f"(v.y) : "memory");
        }
    }
    if (warp == 0) umma::tmem_alloc(reinterpret_cast<uint32_t*>(smem_raw + M::tmem_ptr), XTMEM_COLS);
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_raw + M::tmem_ptr);

    // ---- tile order ----
    const bool rmaj = a.rays && (a.ray_major_dev ? (__ldg(a.ray_major_dev) != 0) : (a.ray_major != 0));
    const int SG4 = (a.S + 3) / 4;
    const int64_t ntiles = rmaj ? ((a.N + 31) / 32) * SG4 : (P + TM - 1) / TM;
    const int64_t nk = (int64_t)blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;   // this CTA's tiles
    auto point_of = [&](int64_t k, int row, bool& on) -> int64_t {
        const int64_t g = (int64_t)blockIdx.x + k * gridDim.x;
        if (rmaj) {
            const int64_t ray = (g / SG4) * 32 + (row & 31);
            const int si = (int)(g % SG4) * 4 + (row >> 5);
            on = k < nk && ray < a.N && si < a.S;
            return on ? ray * a.S + si : 0;
        }
        const int64_t p = g * TM + row;
        on = k < nk && p < P;
        return on ? p : 0;
    };

    if (warp >= XNC * 4) {
        // =============================== producers: hash encode -> shared-memory A tile ===============================
        const int pw = warp - XNC * 4, quarter = pw & 3;
        const int row = quarter * 32 + lane;
        auto load_unit_pos = [&](int64_t k, float* u, bool& on) -> int64_t {
            const int64_t p = point_of(k, row, on);
            u[0] = u[1] = u[2] = 0.5f;
            if (on) {
                if (a.rays) {
                    const float* ry = a.rays + 8 * (p / a.S);
                    const float tv = __ldg(a.t + p);
#pragma unroll
                    for (int c = 0; c < 3; ++c) u[c] = __fadd_rn(__ldg(ry + c), __fmul_rn(__ldg(ry + 3 + c), tv));
                } else {
#pragma unroll
                    for (int c = 0; c < 3; ++c) u[c] = a.x[p * a.xs + c];
                }
                if (a.box6) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) u[c] = world_to_unit1(u[c], __ldg(a.box6 + c), __ldg(a.box6 + 3 + c));
                }
            }
            return p;
        };
        float un[3];
        bool on_n;
        int64_t k = pw >> 2;
        int64_t p_n = load_unit_pos(k, un, on_n);
        for (; k < nk; k += XSLOTS) {
            const float u0 = un[0], u1 = un[1], u2 = un[2];
            const bool on = on_n;
            const int64_t p = p_n;
            p_n = load_unit_pos(k + XSLOTS, un, on_n);                   // next tile's position: in flight during this one
            const int s = (int)(k % XST);
            const uint32_t use = (uint32_t)(k / XST);
            const Tile T = mk_tile(sb + M::stages + (uint32_t)s * M::tile_bytes, E);
            umma::mbar_wait_a(empty0 + 8u * s, (use & 1u) ^ 1u);        // the tensor core has finished with this stage
#pragma unroll 1
            for (int c = 0; c < EC; ++c) {
                uint32_t h[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int l = 4 * c + j;
                    const GridCell g = grid_cell(u0, u1, u2, umma::lds_f32(sb + M::res + 4u * l), a.interp);
                    float2 o;
                    if (l < a.staged_levels) o = level_from_lattice(sb + M::lattice + a.lattice_off[l & 1], (int)umma::lds_f32(sb + M::res + 4u * l) + 1, g);
                    else o = level_from_table(a.table + (((size_t)l << a.log2T) << 1), g, hmask);
                    const __half2 hv = __floats2half2_rn(o.x, o.y);
                    h[j] = *reinterpret_cast<const uint32_t*>(&hv);
                }
                const uint4 q = on ? make_uint4(h[0], h[1], h[2], h[3]) : make_uint4(0u, 0u, 0u, 0u);
                sts128(chunk_addr(T, row, c), q);
                if (enc_out && on) reinterpret_cast<uint4*>(enc_out + p * E)[c] = q;
            }
            umma::fence_async_smem();          // my generic-proxy stores -> visible to the tensor core's operand fetch
            __syncwarp();
            if (lane == 0) mbar_arrive(full0 + 8u * s);
        }
    } else {
        // =============================== consumers: the MLP chain of k_field_fwd_mma ===============================
        const int wg = warp >> 2, row = tid & (TM - 1);
        const bool issuer_warp = (warp & 3) == wg;
        const uint32_t bar_id = 1 + wg;
        const uint32_t done = done0 + 8u * wg;
        const uint32_t H0 = tmem_base + (uint32_t)wg * 128u, H1 = H0 + 64u;     // issuer's view (lane 0)
        const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t h0 = H0 + lane_off, h1 = H1 + lane_off;                  // this thread's lanes
        const Tile Wt0 = mk_tile(sb + wm.t0, E), Wt1 = mk_tile(sb + wm.t1, 64), Whd = mk_tile(sb + wm.hd, 64),
                   Wc0 = mk_tile(sb + wm.c0, 32), Wc1 = mk_tile(sb + wm.c1, 64), Wc2 = mk_tile(sb + wm.c2, 64);
        const Tile Bt0 = mk_tile(sb + wm.bt_t0, 16), Bt1 = mk_tile(sb + wm.bt_t1, 16), Bc0 = mk_tile(sb + wm.bt_c0, 16),
                   Bc1 = mk_tile(sb + wm.bt_c1, 16), One = mk_tile(sb + wm.one, 16);
        auto sync_group = [&]() { umma::fence_before_sync(); umma::bar_sync(bar_id, 128); };
        // layer 1 of local tile k: A = the producers' shared-memory tile; frees the stage when the MMAs complete
        auto issue0 = [&](int64_t k) {
            if (issuer_warp) {
                if (umma::elect_one()) {
                    const int s = (int)(k % XST);
                    umma::mbar_wait_a(full0 + 8u * s, (uint32_t)(k / XST) & 1u);
                    umma::fence_after_sync();
                    mma_fwd_bias(H0, mk_tile(sb + M::stages + (uint32_t)s * M::tile_bytes, E), Wt0, One, Bt0, E);
                    umma::commit_a(done);
                    umma::commit_a(empty0 + 8u * s);
                }
                __syncwarp();
            }
        };
        auto issue = [&](int step) {
            if (issuer_warp) {
                if (umma::elect_one()) {
                    umma::fence_after_sync();
                    switch (step) {
                        case 1: mma_layer_ts(H1, H0, Wt1, 64, 64, &One, &Bt1); break;
                        case 2: mma_layer_ts(H0, H1, Whd, 16, 64, nullptr, nullptr); break;
                        case 3: mma_layer_ts(H1, H0, Wc0, 64, 32, &One, &Bc0); break;
                        case 4: mma_layer_ts(H0, H1, Wc1, 64, 64, &One, &Bc1); break;
                        default: mma_layer_ts(H1, H0, Wc2, 16, 64, nullptr, nullptr); break;
                    }
                    umma::commit_a(done);
                }
                __syncwarp();
            }
        };
        auto load_dir = [&](int64_t k, float* d, bool& on) -> int64_t {
            const int64_t p = point_of(k, row, on);
            const float* dp = dir_of(dirs, dstride, dgroup, p);       // p = 0 (a valid row) when off
            d[0] = __ldg(dp); d[1] = __ldg(dp + 1); d[2] = __ldg(dp + 2);
            return p;
        };
        float b_c2[3];
        load_b_c2(sb + wm.b_c2, b_c2);
        uint32_t phase = 0;
        float dn[3];
        bool on_n;
        int64_t k = wg;
        int64_t p_n = nk > 0 ? load_dir(k, dn, on_n) : 0;
        if (k < nk) issue0(k);
        for (; k < nk; k += XNC) {
            const float cdir[3] = { dn[0], dn[1], dn[2] };
            const bool on = on_n;
            const int64_t p = p_n;
            p_n = load_dir(k + XNC, dn, on_n);
            wait_done(done, phase); epi_hidden_tmem(h0); sync_group(); issue(1);      // -> trunk layer 2
            wait_done(done, phase); epi_hidden_tmem(h1); sync_group(); issue(2);      // -> heads
            wait_done(done, phase);
            float sigma;
            {   // heads: accumulator columns 0..G-1 = geo, 15 = raw sigma; colour input row [sh(16) | geo(G) | 0] -> H0[0,16)
                float v[16], sh[16];
                umma::ld16(h0, v);
                sh16_fast(cdir[0], cdir[1], cdir[2], sh);
                uint32_t r[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) r[j] = pack_h2(sh[2 * j], sh[2 * j + 1]);
                float b[16];
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                    const uint4 u = lds128(sb + wm.b_hd + 16 * qd);
                    b[4 * qd] = __uint_as_float(u.x); b[4 * qd + 1] = __uint_as_float(u.y); b[4 * qd + 2] = __uint_as_float(u.z); b[4 * qd + 3] = __uint_as_float(u.w);
                }
                umma::wait_ld();
                sigma = trunc_exp_fast(v[15] + b[15]);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = j < G ? v[j] + b[j] : 0.0f;
#pragma unroll
                for (int j = 0; j < 8; ++j) r[8 + j] = pack_h2(v[2 * j], v[2 * j + 1]);
                umma::st16(h0, r);
                umma::wait_st();
            }
            sync_group(); issue(3);                                                    // -> colour layer 1
            wait_done(done, phase); epi_hidden_tmem(h1); sync_group(); issue(4);      // -> colour layer 2
            wait_done(done, phase); epi_hidden_tmem(h0); sync_group(); issue(5);      // -> colour out
            wait_done(done, phase);
            // H0 is free again (layer 5 has consumed it): the next tile's first layer runs under this tile's output epilogue
            if (k + XNC < nk) issue0(k + XNC);
            {
                float v[16];
                umma::ld16(h1, v);
                umma::wait_ld();
                if (on) rgb_sigma[p] = make_float4(sigmoid_fast(v[0] + b_c2[0]), sigmoid_fast(v[1] + b_c2[1]), sigmoid_fast(v[2] + b_c2[2]), sigma);
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem_base, XTMEM_COLS);
}

template <int E>
int launch_expert_fwd(acn_ctx* ctx, XArgs a, const float* dirs, int dstride, int dgroup, int64_t P, int G,
                          const acn_field_weights* w, void* enc_out, float* rgb_sigma, const int32_t* range,
                          const int32_t* res_host_or_null, cudaStream_t st) {
    constexpr int NP = XNP;
    using M = XMap<E, XCfg<NP>::stages>;
    // coarse levels staged in shared memory: as many of levels 0, 1 as fit beside the weights and the tile ring
    uint32_t lat_bytes = 0;
    a.staged_levels = 0;
    if (res_host_or_null && a.box6) {      // only with the world->unit clamp, which keeps every corner inside the lattice
        for (int l = 0; l < 2 && l < a.L; ++l) {
            const int64_t R1 = (int64_t)res_host_or_null[l] + 1;
            const int64_t bytes = ((R1 * R1 * R1 * 8 + 127) / 128) * 128;
            if (R1 < 2 || (int64_t)M::lattice + lat_bytes + bytes > (int64_t)ctx->max_smem_optin) break;
            a.lattice_off[l] = lat_bytes;
            lat_bytes += (uint32_t)bytes;
            a.staged_levels = l + 1;
        }
    }
    uint32_t smem = M::lattice + lat_bytes;
    ACN_REQUIRE((int)smem <= ctx->max_smem_optin, ACN_EUNSUPPORTED, "acn_render_expert_fwd: needs %u B shared memory", smem);
    const int64_t tiles_a = (P + TM - 1) / TM;
    const int64_t tiles_b = a.rays ? ((a.N + 31) / 32) * ((a.S + 3) / 4) : 0;
    int64_t grid = tiles_a > tiles_b ? tiles_a : tiles_b;
    if (grid > ctx->sm_count) grid = ctx->sm_count;
    if (grid < 1) grid = 1;
    ACN_CUDA(cudaFuncSetAttribute(k_expert_fwd<E, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_expert_fwd<E, NP><<<(int)grid, XCfg<NP>::threads, smem, st>>>(a, dirs, dstride, dgroup, P, G, *w, (__half*)enc_out,
                                                                             (float4*)rgb_sigma, range);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

}  // namespace

extern "C" int acn_render_expert_fwd(acn_ctx* ctx, const float* x_or_null, int x_stride, const float* rays8_or_null,
                                     const float* t_vals_or_null, int64_t P, int S, int ray_major,
                                     const int32_t* ray_major_dev_or_null, const int32_t* range_or_null,
                                     const float* box6_or_null, const float* table, int L, int F, int log2T,
                                     const int32_t* res, const int32_t* res_host_or_null, int interp, const float* dirs,
                                     int dirs_stride, int dirs_group, int H, int G, int C, const acn_field_weights* w,
                                     void* enc_f16_out_or_null, float* rgb_sigma, acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    const char* fn = "acn_render_expert_fwd";
    ACN_REQUIRE(P >= 0, ACN_EINVAL, "%s: negative P", fn);
    ACN_REQUIRE(F == 2 && (L == 8 || L == 16), ACN_EUNSUPPORTED, "%s: built for F = 2 and 8 or 16 levels (got L=%d, F=%d)", fn, L, F);
    ACN_REQUIRE(log2T >= 1 && log2T <= 24 && res, ACN_EINVAL, "%s: bad table size / res table", fn);
    ACN_REQUIRE(interp == ACN_INTERP_LINEAR || interp == ACN_INTERP_SMOOTHSTEP, ACN_EUNSUPPORTED, "%s: interpolation must be Linear or Smoothstep", fn);
    ACN_REQUIRE(H == 64 && C == 64, ACN_EUNSUPPORTED, "%s: hidden widths must be 64 (got H=%d, C=%d)", fn, H, C);
    ACN_REQUIRE(G >= 1 && G <= 15, ACN_EUNSUPPORTED, "%s: geo_feat_dim %d outside [1,15]", fn, G);
    ACN_REQUIRE(w, ACN_EINVAL, "%s: null weights", fn);
    for (int i = 0; i < 14; ++i) ACN_REQUIRE(w->p[i] != nullptr, ACN_EINVAL, "%s: weight pointer %d is null", fn, i);
    ACN_REQUIRE(dirs_stride >= 3 && dirs_group >= 1, ACN_EINVAL, "%s: bad dirs stride/group", fn);
    if (P == 0) return ACN_OK;
    const bool from_rays = rays8_or_null != nullptr;
    ACN_REQUIRE(from_rays ? (t_vals_or_null && S >= 1 && P % S == 0) : (x_or_null && x_stride >= 3), ACN_EINVAL,
                "%s: give either x (P,>=3) or rays8 + t_vals with P = N*S", fn);
    ACN_REQUIRE(!range_or_null || (!from_rays && dirs_group == 1), ACN_EINVAL, "%s: a row range needs explicit positions and per-point directions", fn);
    ACN_REQUIRE(table && dirs && rgb_sigma, ACN_EINVAL, "%s: null buffer", fn);
    ACN_REQUIRE((((uintptr_t)table | (uintptr_t)rgb_sigma | (uintptr_t)enc_f16_out_or_null) & 15) == 0, ACN_EINVAL,
                "%s: table / rgb_sigma / enc misaligned", fn);
    XArgs a{};
    a.x = from_rays ? nullptr : x_or_null; a.xs = x_stride;
    a.rays = rays8_or_null; a.t = t_vals_or_null; a.S = from_rays ? S : 1; a.N = from_rays ? P / S : 0;
    a.ray_major = ray_major; a.ray_major_dev = ray_major_dev_or_null;
    a.box6 = box6_or_null; a.table = table; a.L = L; a.log2T = log2T; a.res = res; a.interp = interp;
    cudaStream_t st = (cudaStream_t)stream;
    if (L == 8) return launch_expert_fwd<16>(ctx, a, dirs, dirs_stride, dirs_group, P, G, w, enc_f16_out_or_null, rgb_sigma, range_or_null, res_host_or_null, st);
    return launch_expert_fwd<32>(ctx, a, dirs, dirs_stride, dirs_group, P, G, w, enc_f16_out_or_null, rgb_sigma, range_or_null, res_host_or_null, st);
}
