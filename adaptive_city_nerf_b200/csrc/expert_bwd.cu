// Fused per-expert backward, warp-specialised (acn_render_expert_bwd; replaces autograd through
// nerfs/ray_rendering.py:317-325 -> models/inr/meta_ngp.py:226-241 -> models/encodings.py:331-381): the tcgen05 MLP
// backward of field_mma.cu and the hash-table gradient scatter of hashgrid.cu in ONE persistent kernel in which the two
// do not share warps.  The (P, L*F) gradient of the encoding never exists in HBM.
//
//   * 8 CHAIN warps (2 tile slots x 128 threads, one thread per point row) run the layer chain of k_field_bwd_mma:
//     forward recompute, dgrad, weight gradients into TMEM-resident accumulators.  The LAST dgrad of a tile (d enc) is
//     issued into a dedicated 32-column TMEM window per slot instead of the slot's D window, and the chain moves on to
//     its next tile without ever reading it.
//   * 8 SCATTER warps (2 per TMEM lane quarter; warp = 32 consecutive points) wait for that MMA's commit, read their
//     columns of the window (tcgen05.ld), release it, and add w_corner * g to the 8 corner rows of the table gradient at
//     each of their levels -- the arithmetic of k_hashgrid_bwd, with its pair REDs and the run-merging of the coarse
//     levels.  Levels are dealt round-robin (part p: p, p + 2, ...) so both parts carry one of the costlier run-merged coarse levels.
//
// Measured on the bench batch (2^24 samples, profiles/r02_v6_fused_bwd_variants.txt): 8.2 - 8.4 ms, against 10.0 ms for the
// single-role kernel (k_field_bwd_mma<.., SCAT>: every MLP thread scatters its own columns between its epilogues, whose
// chain time and RED time add up) and 9.8 / 9.9 ms with 16 / 12 scatter warps at 64 / 80 registers: a RED holds its
// payload and address registers until the memory system has taken it, so what counts is registers per scatter warp
// (144 here: eight distinct payload quads in flight) more than the number of warps.  Registers: 512 threads launch with
// 128 each; setmaxnreg moves registers inside that CTA pool only (an .inc that counts on the SM's unallocated registers
// never completes): the chain warps hand back 16 (-> 112), the scatter warps take them (-> 144).
#include "field_mma.cuh"

namespace {

constexpr int YCW = BNS * 4;                 // chain warps
constexpr int YSW = 8;                       // scatter warps
constexpr int YPARTS = YSW / 4;              // ... per TMEM lane quarter: each takes every YPARTS-th level
constexpr int YTHREADS = (YCW + YSW) * 32;   // 512 threads x 128 registers = the whole register file
constexpr uint32_t COL_DENC = 448;           // per-slot 32-column window the last dgrad writes d_enc into (448 .. 511)
constexpr int DEDUP_LEVELS = 2;           // run-merged coarse levels: 2 measured best (8.71 ms; 3: 8.71, 4: 8.81), one per scatter part

static_assert(DEDUP_LEVELS % YPARTS == 0, "every scatter part takes the same number of run-merged levels");

template <int E> struct YMap {
    static constexpr uint32_t slots = 0;
    static constexpr uint32_t w = BNS * SlotMap<E>::bytes;
    static constexpr uint32_t bars = (w + wmap(E).end + 127u) & ~127u;     // per slot: done_d, done_w, denc_full, denc_free
    static constexpr uint32_t tmem_ptr = bars + BNS * 4 * 8u;
    static constexpr uint32_t res = tmem_ptr + 16u;                         // float resolution per level (16)
    static constexpr uint32_t bytes = res + 64u;
};

__device__ __forceinline__ void mbar_arrive_a(uint32_t addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(addr) : "memory");
}
__device__ __forceinline__ void ld2(uint32_t taddr, float* v) {
    uint32_t r0, r1;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr) : "memory");
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1);
}
template <int E>
__global__ void __launch_bounds__(YTHREADS, 1) k_expert_bwd(
    const __half* __restrict__ enc, const float* __restrict__ dirs, int dstride, int dgroup, int64_t P, int G,
    acn_field_weights w, const float4* __restrict__ d_rgb_sigma, const unsigned int* __restrict__ absmax_bits,
    acn_field_grads g, ScatterArgs sc, const int32_t* __restrict__ range)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    if (range) {            // rows [range[0], range[1]) only (an expert's bucket): P was just the launch's upper bound
        const int64_t r0 = __ldg(range);
        P = __ldg(range + 1) - r0;
        enc += r0 * E; dirs += r0 * dstride; d_rgb_sigma += r0;
        if (sc.x) sc.x += r0 * sc.xs;
    }
    if (P <= 0) return;     // an empty bucket (the host never reads the counts): leave before staging weights / allocating TMEM
    using M = YMap<E>;
    using SM = SlotMap<E>;
    constexpr WMap wm = wmap(E);
    constexpr int EC = E / 8;            // 16-byte chunks per encoding row
    const uint32_t sb = umma::smem_u32(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    stage_weights<E>(w, G, sb + M::w);
    for (int i = tid; i < BNS * 4 * TM; i += blockDim.x) {      // the ones chunk (chunk 8) of every hidden tile
        const int r = i & (TM - 1), t = (i >> 7) & 3, sl = i >> 9;
        const Tile ht = mk_tile(sb + M::slots + (uint32_t)sl * SM::bytes + SM::c2 + (uint32_t)t * (TM * HW * 2), HW);
        sts128(chunk_addr(ht, r, 8), make_uint4(0x00003C00u, 0u, 0u, 0u));
    }
    if (tid == 0) {
        for (int i = 0; i < BNS; ++i) {
            umma::mbar_init_a(sb + M::bars + 32 * i, 1);                 // done_d: the chain's MMAs of a step
            umma::mbar_init_a(sb + M::bars + 32 * i + 8, 1);             // done_w: the weight-gradient MMAs of a step
            umma::mbar_init_a(sb + M::bars + 32 * i + 16, 1);            // denc_full: the last dgrad has landed in the window
            umma::mbar_init_a(sb + M::bars + 32 * i + 24, YSW * 32);     // denc_free: every scatter thread has read its columns
        }
        umma::fence_mbar_init();
    }
    if (tid < 16) umma::sts_f32(sb + M::res + 4u * tid, tid < sc.L ? (float)__ldg(sc.res + tid) : 1.0f);
    if (warp == 0) umma::tmem_alloc(reinterpret_cast<uint32_t*>(smem_raw + M::tmem_ptr), BWD_TMEM_COLS);
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_raw + M::tmem_ptr);
    if (warp < 4) {   // zero the persistent accumulators (lanes 32*warp.., every accumulator column)
        for (uint32_t c = COL_ACC0; c < COL_ACC1; c += 16) umma::st16_zero(tmem_base + ((uint32_t)(warp * 32) << 16) + c);
        umma::wait_st();
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();

    const int64_t ntiles = (P + TM - 1) / TM;
    const int64_t tile_stride = (int64_t)gridDim.x * BNS;
    const float scale = grad_scale_from_max(__uint_as_float(__ldg(absmax_bits)));
    const float inv_scale = 1.0f / scale;

    if (warp >= YCW) {
        // ======================================= scatter warps =======================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 144;");
        constexpr int L = E / 2, LPP = L / YPARTS, NCOL = 2 * LPP;      // levels / TMEM columns per scatter warp
        const int quad = warp & 3, part = (warp - YCW) >> 2;
        const uint32_t hmask = (1u << sc.log2T) - 1u;
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        uint32_t ph_full[BNS];
#pragma unroll
        for (int sl = 0; sl < BNS; ++sl) ph_full[sl] = 0;
        for (int64_t it = 0;; ++it) {
            bool any = false;
#pragma unroll
            for (int sl = 0; sl < BNS; ++sl) {
                const int64_t stile = (int64_t)blockIdx.x * BNS + sl + it * tile_stride;
                if (stile >= ntiles) continue;
                any = true;
                const int64_t sp = stile * TM + quad * 32 + lane;
                const bool son = sp < P;
                float upos[3] = { 0.5f, 0.5f, 0.5f };
                if (son) {   // the point's unit-cube position; the loads are in flight while the tile's layers run
                    if (sc.rays) {
                        const float* ry = sc.rays + 8 * (sp / sc.S);
                        const float tv = __ldg(sc.t + sp);
#pragma unroll
                        for (int c = 0; c < 3; ++c) upos[c] = __fadd_rn(__ldg(ry + c), __fmul_rn(__ldg(ry + 3 + c), tv));
                    } else {
#pragma unroll
                        for (int c = 0; c < 3; ++c) upos[c] = __ldg(sc.x + sp * sc.xs + c);
                    }
                    if (sc.box6) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) upos[c] = world_to_unit1(upos[c], __ldg(sc.box6 + c), __ldg(sc.box6 + 3 + c));
                    }
                }
                const uint32_t bfull = sb + M::bars + 32u * sl + 16u, bfree = bfull + 8u;
                wait_done(bfull, ph_full[sl]);
                // levels are dealt round-robin over the parts (part p: p, p + YPARTS, ...): the run-merged coarse levels cost about
                // twice a fine one, and with a contiguous split one part had all of them and set the pace for everybody
                float v[NCOL];
                const uint32_t win = tmem_base + COL_DENC + 32u * sl + lane_off;
#pragma unroll
                for (int j = 0; j < LPP; ++j) ld2(win + 2u * (uint32_t)(part + j * YPARTS), v + 2 * j);
                umma::wait_ld();
                umma::fence_before_sync();
                mbar_arrive_a(bfree);          // the chain issuer may overwrite the window with the slot's next tile
                // j = 0: this part's run-merged coarse level; j >= 1: fine levels, one ROLLED loop (the unrolled form was 7 copies of
                // the cell arithmetic and its REDs: instruction-cache footprint matters with two roles running different code)
                auto one_level = [&](int l, float gx, float gy, bool dedup) {
                    const bool act = son && !(gx == 0.0f && gy == 0.0f);      // fully occluded samples scatter nothing
                    if (!dedup && !act) return;
                    const GridCell c = grid_cell(upos[0], upos[1], upos[2], umma::lds_f32(sb + M::res + 4u * l), sc.interp);
                    const float wx[2] = { 1.0f - c.wx, c.wx }, wy[2] = { 1.0f - c.wy, c.wy }, wz[2] = { 1.0f - c.wz, c.wz };
                    float2 acc[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float wk = act ? wz[k & 1] * wy[(k >> 1) & 1] * wx[(k >> 2) & 1] : 0.0f;
                        acc[k] = make_float2(gx * wk, gy * wk);
                    }
                    float2* lt = reinterpret_cast<float2*>(sc.dtable) + ((size_t)l << sc.log2T);
                    if (dedup) {
                        // Coarse levels (a ray stays in one cell for several samples and the lanes of a warp are 32 consecutive
                        // points): the lanes of a run of equal cells add their contributions with a segmented warp scan and only
                        // the run's last lane issues the REDs.  Cells have < 2^10 cells per axis; an idle lane is its own run.
                        const uint32_t key = act ? ((c.x0 & 1023u) | ((c.y0 & 1023u) << 10) | ((c.z0 & 1023u) << 20)) : (0x80000000u | (uint32_t)lane);
                        const uint32_t kprev = __shfl_up_sync(0xffffffffu, key, 1);
                        const bool head = lane == 0 || kprev != key;
                        bool closed = head;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const bool cu = __shfl_up_sync(0xffffffffu, (int)closed, d) != 0;
                            const bool take = lane >= d && !closed;
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                const float tx = __shfl_up_sync(0xffffffffu, acc[k].x, d), ty = __shfl_up_sync(0xffffffffu, acc[k].y, d);
                                if (take) { acc[k].x += tx; acc[k].y += ty; }
                            }
                            if (lane >= d) closed = closed || cu;
                        }
                        const bool hnext = __shfl_down_sync(0xffffffffu, (int)head, 1) != 0;
                        if (act && (lane == 31 || hnext)) scatter_cell_f2(lt, c.x0, c.y0, c.z0, hmask, acc);
                    } else {
                        scatter_cell_f2(lt, c.x0, c.y0, c.z0, hmask, acc);
                    }
                };
                static_assert(DEDUP_LEVELS == YPARTS, "one run-merged level per part: j = 0");
                one_level(part, v[0] * inv_scale, v[1] * inv_scale, true);
#pragma unroll 1
                for (int j = 1; j < LPP; ++j) {
                    float gx = v[2], gy = v[3];                 // v[2j], v[2j+1] by a select chain: v stays in registers
#pragma unroll
                    for (int q = 2; q < LPP; ++q) { gx = j == q ? v[2 * q] : gx; gy = j == q ? v[2 * q + 1] : gy; }
                    one_level(part + j * YPARTS, gx * inv_scale, gy * inv_scale, false);
                }
            }
            if (!any) break;
        }
    } else {
        // ======================================= chain warps =======================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 112;");
        const int slot = warp >> 2;                    // 4 warps per slot, thread = row
        const int row = (warp & 3) * 32 + lane;
        const bool issuer_warp = (warp & 3) == slot;                   // slot 0 -> warp 0, slot 1 -> warp 5
        const bool wgrad_warp = (warp & 3) == ((slot + 2) & 3);        // slot 0 -> warp 2, slot 1 -> warp 7: other sub-partitions
        const uint32_t bar_id = 1 + slot;
        const uint32_t sbase = sb + M::slots + (uint32_t)slot * SM::bytes;
        const Tile Txe = mk_tile(sbase + SM::xe, E), Th1 = mk_tile(sbase + SM::h1, HW), Th2 = mk_tile(sbase + SM::h2, HW),
                   Tcin = mk_tile(sbase + SM::cin, 32), Tc1 = mk_tile(sbase + SM::c1, HW), Tc2 = mk_tile(sbase + SM::c2, HW),
                   Tdrr = mk_tile(sbase + SM::dg, 16), Tghd = mk_tile(sbase + SM::dg, 16);
        const uint32_t wb = sb + M::w;
        const Tile Wt0 = mk_tile(wb + wm.t0, E), Wt1 = mk_tile(wb + wm.t1, 64), Whd = mk_tile(wb + wm.hd, 64),
                   Wc0 = mk_tile(wb + wm.c0, 32), Wc1 = mk_tile(wb + wm.c1, 64), Wc2 = mk_tile(wb + wm.c2, 64);
        const Tile Bt0 = mk_tile(wb + wm.bt_t0, 16), Bt1 = mk_tile(wb + wm.bt_t1, 16), Bc0 = mk_tile(wb + wm.bt_c0, 16),
                   Bc1 = mk_tile(wb + wm.bt_c1, 16), One = mk_tile(wb + wm.one, 16);
        const uint32_t done_d = sb + M::bars + 32u * slot, done_w = done_d + 8u, denc_full = done_d + 16u, denc_free = done_d + 24u;
        const uint32_t dwin = tmem_base + (uint32_t)slot * 64u;
        const uint32_t tmem_d = dwin + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t dencwin = tmem_base + COL_DENC + 32u * (uint32_t)slot;
        uint32_t ph_free = 1;        // parity the issuer waits for on denc_free (a fresh barrier passes parity 1)

        auto issue = [&](int step) {
            if (issuer_warp) {
                if (umma::elect_one()) {
                    umma::fence_after_sync();
                    switch (step) {
                        case 0: mma_fwd_bias(dwin, Txe, Wt0, One, Bt0, E); break;
                        case 1: mma_fwd_bias(dwin, Th1, Wt1, One, Bt1, 64); break;
                        case 2: mma_fwd(dwin, Th2, Whd, 16, 64); break;
                        case 3: mma_fwd_bias(dwin, Tcin, Wc0, One, Bc0, 32); break;
                        case 4: mma_fwd_bias(dwin, Tc1, Wc1, One, Bc1, 64); break;
                        case 5: mma_fwd(dwin, Tc2, Wc2, 16, 64); break;
                        case 6: mma_dgrad(dwin, Tdrr, Wc2, 64, 16); break;          // colour out: D = drr W_c2
                        case 7: mma_dgrad(dwin, Tc2, Wc1, 64, 64); break;           // colour layer 2 (g lives in the c2 tile now)
                        case 8: mma_dgrad(dwin, Tc1, Wc0, 32, 64); break;           // colour layer 1
                        case 9: mma_dgrad(dwin, Tghd, Whd, 64, 16); break;          // heads
                        case 10: mma_dgrad(dwin, Th2, Wt1, 64, 64); break;          // trunk layer 2
                        default:   // trunk layer 1: d enc, into the scatter warps' window once they have read the previous tile's
                            umma::mbar_wait_a(denc_free, ph_free);
                            umma::fence_after_sync();
                            mma_dgrad(dencwin, Th1, Wt0, E, 64);
                            break;
                    }
                    // step 11 is committed to denc_full only: nobody in the chain reads d enc, and this thread's next
                    // commit to done_d (the next tile's first layer) covers it, MMAs of one thread completing in order
                    umma::commit_a(step == 11 ? denc_full : done_d);
                }
                __syncwarp();
            }
            if (wgrad_warp && step >= 6) {
                if (umma::elect_one()) {
                    umma::fence_after_sync();
                    switch (step) {
                        case 6: mma_over_points(tmem_base + COL_WC2T, Tc2, Tdrr, 16, 128); break;   // [dW^T ; db] = [c2 | 1]^T drr
                        case 7: mma_over_points(tmem_base + COL_WC1, Tc2, Tc1, HW); break;          // [dW | db] = g^T [c1 | 1]
                        case 8: mma_over_points(tmem_base + COL_WC0, Tc1, Tcin, 32); break;         // column 31 of cin is the ones column
                        case 9: mma_over_points(tmem_base + COL_WHDT, Th2, Tghd, 16, 128); break;   // [dW^T ; db] = [h2 | 1]^T ghd
                        case 10: mma_over_points(tmem_base + COL_WT1, Th2, Th1, HW); break;
                        default:   // the encoding tile has no spare column: separate G^T 1
                            mma_over_points(tmem_base + COL_WT0, Th1, Txe, E);
                            mma_over_points(tmem_base + COL_BT0, Th1, One, 8);
                            break;
                    }
                    umma::commit_a(done_w);
                }
                __syncwarp();
            }
        };
        auto load_inputs = [&](int64_t tile, uint4* q, float* dir, float4& dy) {
            const int64_t p = tile * TM + row;
            const bool on = tile < ntiles && p < P;
#pragma unroll
            for (int c = 0; c < EC; ++c) q[c] = on ? __ldg(reinterpret_cast<const uint4*>(enc + p * E) + c) : make_uint4(0u, 0u, 0u, 0u);
            dir[0] = 0.f; dir[1] = 0.f; dir[2] = 1.f;
            dy = make_float4(0.f, 0.f, 0.f, 0.f);
            if (on) {
                const float* dp = dir_of(dirs, dstride, dgroup, p);
                dir[0] = __ldg(dp); dir[1] = __ldg(dp + 1); dir[2] = __ldg(dp + 2);
                dy = __ldg(d_rgb_sigma + p);
            }
        };
        auto hidden = [&](const Tile& dst) { epi_hidden32(tmem_d, 0, dst, row); epi_hidden32(tmem_d, 32, dst, row); };
        // backward epilogue of a hidden layer: D -> fp16 -> * [act > 0] (registers), then -- once the weight-gradient MMAs that
        // still read the activation tile have completed -- overwrite it in place with the gradient tile
        uint32_t ph_d = 0, ph_w = 0;
        auto masked = [&](const Tile& act) {
            uint4 o[8];
            wait_done(done_d, ph_d);
            epi_mask32_load(tmem_d, 0, act, row, o);
            epi_mask32_load(tmem_d, 32, act, row, o + 4);
            wait_done(done_w, ph_w);
            epi_mask32_store(act, row, 0, o);
            epi_mask32_store(act, row, 32, o + 4);
        };

        uint4 encq[EC];
        float dir[3];
        float4 dy;
        int64_t tile = (int64_t)blockIdx.x * BNS + slot;
        load_inputs(tile, encq, dir, dy);
        for (; tile < ntiles; tile += tile_stride) {
            // ---------------- forward recompute (same arithmetic as k_field_fwd_mma) ----------------
#pragma unroll
            for (int c = 0; c < EC; ++c) sts128(chunk_addr(Txe, row, c), encq[c]);
            group_sync(bar_id, 128); issue(0);
            const float cdir[3] = { dir[0], dir[1], dir[2] };
            const float4 cdy = dy;
            load_inputs(tile + tile_stride, encq, dir, dy);        // prefetch: lands while this tile runs
            wait_done(done_d, ph_d); hidden(Th1); group_sync(bar_id, 128); issue(1);
            wait_done(done_d, ph_d); hidden(Th2); group_sync(bar_id, 128); issue(2);
            wait_done(done_d, ph_d);
            const float sig_raw = epi_heads_geo(tmem_d, wb + wm.b_hd, G, Tcin, row, 1.0f);
            epi_heads_sh(cdir, Tcin, row);
            group_sync(bar_id, 128); issue(3);
            wait_done(done_d, ph_d); hidden(Tc1); group_sync(bar_id, 128); issue(4);
            wait_done(done_d, ph_d); hidden(Tc2); group_sync(bar_id, 128); issue(5);
            wait_done(done_d, ph_d);
            float d_sig;
            {   // output gradients (scaled): d rgb_raw = dy * y (1 - y); d sigma_raw = dy * exp(clamp(sigma_raw))
                float v[16], drr[16], bc2[3];
                umma::ld16(tmem_d, v);
                load_b_c2(wb + wm.b_c2, bc2);
                umma::wait_ld();
                const float y0 = sigmoid_fast(v[0] + bc2[0]), y1 = sigmoid_fast(v[1] + bc2[1]), y2 = sigmoid_fast(v[2] + bc2[2]);
#pragma unroll
                for (int j = 0; j < 16; ++j) drr[j] = 0.0f;
                drr[0] = cdy.x * scale * y0 * (1.0f - y0); drr[1] = cdy.y * scale * y1 * (1.0f - y1); drr[2] = cdy.z * scale * y2 * (1.0f - y2);
                d_sig = cdy.w * scale * trunc_exp_fast(sig_raw);
                sts128(chunk_addr(Tdrr, row, 0), pack8(drr));
                sts128(chunk_addr(Tdrr, row, 1), pack8(drr + 8));
            }
            group_sync(bar_id, 128); issue(6);
            // ---------------- backward ----------------
            masked(Tc2); group_sync(bar_id, 128); issue(7);
            masked(Tc1); group_sync(bar_id, 128); issue(8);
            wait_done(done_d, ph_d);
            {   // d cin = [d sh | d geo | 0] -> heads gradient row [d geo | 0 | d sigma_raw @15]
                float v[16];
                umma::ld16(tmem_d + 16, v);
                umma::wait_ld();
#pragma unroll
                for (int j = 0; j < 15; ++j) v[j] = j < G ? v[j] : 0.0f;
                v[15] = d_sig;
                sts128(chunk_addr(Tghd, row, 0), pack8(v));
                sts128(chunk_addr(Tghd, row, 1), pack8(v + 8));
            }
            wait_done(done_w, ph_w);
            group_sync(bar_id, 128); issue(9);
            masked(Th2); group_sync(bar_id, 128); issue(10);
            masked(Th1); group_sync(bar_id, 128); issue(11);
            ph_free ^= 1u;
            wait_done(done_w, ph_w);     // the xe / h1 tiles are free again only now
        }
        // ---------------- add the TMEM-resident weight gradients to global memory ----------------
        umma::fence_before_sync();
        umma::bar_sync(3, YCW * 32);           // both slots' chains are done with the accumulators
        umma::fence_after_sync();
        if (warp < 4) {
            // the wgrad MMAs were all waited for (done_w) by their slots; the last commit of each slot's issuer covers them
            const uint32_t tmem_row = tmem_base + ((uint32_t)(warp * 32) << 16);
            const int t = (lane < 16) ? warp * 16 + lane : 1 << 20;     // row of this lane in an M=64 accumulator (or none)
            const int u = tid;                                           // row of this lane in an M=128 accumulator
            auto flush = [&](uint32_t col, int ncols, float* dst, int ld_col, int nvalid) {
                for (int q = 0; q < ncols / 16; ++q) {
                    float v[16];
                    umma::ld16(tmem_row + col + q * 16, v);     // warp-collective: every lane loads
                    umma::wait_ld();
                    if (dst) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int n = q * 16 + j;
                            if (n < nvalid && v[j] != 0.0f) atomicAdd(dst + (size_t)n * ld_col, v[j] * inv_scale);
                        }
                    }
                }
            };
            auto at = [&](float* base, bool valid, int off) { return (base && valid) ? base + off : (float*)nullptr; };
            const int CIN = G + 16;
            flush(COL_WC1, 64, at(g.p[10], t < 64, t * 64), 1, 64);
            flush(COL_WC1 + 64, 16, at(g.p[11], t < 64, t), 1, 1);
            flush(COL_WT1, 64, at(g.p[2], t < 64, t * 64), 1, 64);
            flush(COL_WT1 + 64, 16, at(g.p[3], t < 64, t), 1, 1);
            flush(COL_WT0, 64, at(g.p[0], t < 64, t * E), 1, E);
            flush(COL_BT0, 16, at(g.p[1], t < 64, t), 1, 1);
            flush(COL_WC0, 16, at(g.p[8], t < 64, t * CIN + G), 1, 16);
            {
                float v[16];
                umma::ld16(tmem_row + COL_WC0 + 16, v);
                umma::wait_ld();
                if (t < 64) {
#pragma unroll
                    for (int j = 0; j < 15; ++j) if (g.p[8] && j < G && v[j] != 0.0f) atomicAdd(g.p[8] + t * CIN + j, v[j] * inv_scale);
                    if (g.p[9] && v[15] != 0.0f) atomicAdd(g.p[9] + t, v[15] * inv_scale);
                }
            }
            flush(COL_WC2T, 16, u < 64 ? at(g.p[12], true, u) : (u == 64 ? g.p[13] : (float*)nullptr), u < 64 ? 64 : 1, 3);
            {
                float v[16];
                umma::ld16(tmem_row + COL_WHDT, v);
                umma::wait_ld();
                if (u < 64) {
#pragma unroll
                    for (int j = 0; j < 15; ++j) if (g.p[6] && j < G && v[j] != 0.0f) atomicAdd(g.p[6] + j * 64 + u, v[j] * inv_scale);
                    if (g.p[4] && v[15] != 0.0f) atomicAdd(g.p[4] + u, v[15] * inv_scale);
                } else if (u == 64) {
#pragma unroll
                    for (int j = 0; j < 15; ++j) if (g.p[7] && j < G && v[j] != 0.0f) atomicAdd(g.p[7] + j, v[j] * inv_scale);
                    if (g.p[5] && v[15] != 0.0f) atomicAdd(g.p[5], v[15] * inv_scale);
                }
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem_base, BWD_TMEM_COLS);
}

template <int E>
int launch_expert_bwd(acn_ctx* ctx, const void* enc, const float* dirs, int dirs_stride, int dirs_group, int64_t P, int G,
                      const acn_field_weights* w, const float* d_rgb_sigma, const unsigned int* absmax, const acn_field_grads* g,
                      const ScatterArgs& sc, const int32_t* range, cudaStream_t st) {
    constexpr uint32_t smem = YMap<E>::bytes;
    ACN_REQUIRE((int)smem <= ctx->max_smem_optin, ACN_EUNSUPPORTED, "acn_render_expert_bwd: needs %u B shared memory", smem);
    const int64_t ntiles = (P + TM - 1) / TM;
    int64_t grid = (ntiles + BNS - 1) / BNS;
    if (grid > ctx->sm_count) grid = ctx->sm_count;
    ACN_CUDA(cudaFuncSetAttribute(k_expert_bwd<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_expert_bwd<E><<<(int)grid, YTHREADS, smem, st>>>((const __half*)enc, dirs, dirs_stride, dirs_group, P, G, *w,
                                                       (const float4*)d_rgb_sigma, absmax, *g, sc, range);
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

}  // namespace

// Fully fused backward of one expert on a batch of points (SURVEY 8b acn_render_expert_bwd).
extern "C" int acn_render_expert_bwd(acn_ctx* ctx, const float* x_or_null, int x_stride, const float* rays8_or_null,
                                     const float* t_vals_or_null, int64_t P, int S, const int32_t* range_or_null,
                                     const float* box6_or_null, int L, int F,
                                     int log2T, const int32_t* res, int interp, const void* enc_f16, const float* dirs,
                                     int dirs_stride, int dirs_group, int H, int G, int C, const acn_field_weights* w,
                                     const float* d_rgb_sigma, const acn_field_grads* g, float* dtable, acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    const char* fn = "acn_render_expert_bwd";
    ACN_REQUIRE(P >= 0, ACN_EINVAL, "%s: negative P", fn);
    ACN_REQUIRE(F == 2 && (L == 8 || L == 16), ACN_EUNSUPPORTED, "%s: built for F = 2 and 8 or 16 levels (got L=%d, F=%d)", fn, L, F);
    ACN_REQUIRE(log2T >= 1 && log2T <= 24 && res, ACN_EINVAL, "%s: bad table size / res table", fn);
    ACN_REQUIRE(interp == ACN_INTERP_LINEAR || interp == ACN_INTERP_SMOOTHSTEP, ACN_EUNSUPPORTED, "%s: interpolation must be Linear or Smoothstep", fn);
    ACN_REQUIRE(w && g, ACN_EINVAL, "%s: null weights / grads", fn);
    for (int i = 0; i < 14; ++i) ACN_REQUIRE(w->p[i] != nullptr, ACN_EINVAL, "%s: weight pointer %d is null", fn, i);
    ACN_REQUIRE(dirs_stride >= 3 && dirs_group >= 1, ACN_EINVAL, "%s: bad dirs stride/group", fn);
    if (P == 0) return ACN_OK;
    const bool from_rays = rays8_or_null != nullptr;
    ACN_REQUIRE(from_rays ? (t_vals_or_null && S >= 1 && P % S == 0) : (x_or_null && x_stride >= 3), ACN_EINVAL,
                "%s: give either x (P,>=3) or rays8 + t_vals with P = N*S", fn);
    ACN_REQUIRE(!range_or_null || (!from_rays && dirs_group == 1), ACN_EINVAL, "%s: a row range needs explicit positions and per-point directions", fn);
    ACN_REQUIRE(enc_f16 && dirs && d_rgb_sigma && dtable, ACN_EINVAL, "%s: null buffer", fn);
    ACN_REQUIRE((((uintptr_t)d_rgb_sigma | (uintptr_t)dtable) & 15) == 0, ACN_EINVAL, "%s: d_rgb_sigma / dtable misaligned", fn);
    const int E = L * F;
    int rc = check_dims(fn, ACN_F16, E, H, G, C, enc_f16);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned int* slot = nullptr;
    rc = absmax_word(ctx, fn, d_rgb_sigma, P, range_or_null, st, &slot);      // loss scale: max |dL/dy| over the batch
    if (rc) return rc;
    const ScatterArgs sc{ from_rays ? nullptr : x_or_null, x_stride, rays8_or_null, t_vals_or_null, S, box6_or_null, dtable, L, log2T, res, interp };
    if (E == 16) return launch_expert_bwd<16>(ctx, enc_f16, dirs, dirs_stride, dirs_group, P, G, w, d_rgb_sigma, slot, g, sc, range_or_null, st);
    return launch_expert_bwd<32>(ctx, enc_f16, dirs, dirs_stride, dirs_group, P, G, w, d_rgb_sigma, slot, g, sc, range_or_null, st);
}
