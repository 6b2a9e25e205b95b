// Stage 3 under autocast(fp16): the expert's field MLPs on tcgen05 tensor cores, forward and backward.
//
// Reference numerics: models/metamodule/metamodule.py:150-155 under torch.autocast -- fp16 GEMM operands, fp32
// accumulation, fp32 bias + ReLU, fp16 activations into the next GEMM.  Here the bias (as a hi + lo fp16 pair, i.e. to
// fp32 accuracy) is added inside the fp32 accumulator by one extra K step, so a hidden activation is
// fp16(relu(sum + b)): one rounding where the reference has two (GEMM result, then the next layer's input cast).
//
// Both kernels are persistent (one CTA per SM).  Tiles are 128 points (MMA M = 128); thread r of a tile's warpgroup owns
// point r: lane r of the tile's TMEM window (and, in the backward, row r of its shared-memory tiles, canonical
// no-swizzle UMMA layout, umma.cuh).  All six weight matrices stay resident in shared memory as B operands.  The layers
// of a tile are serially dependent (MMA -> epilogue -> MMA ...), so several tiles run out of phase per SM, with no
// __syncthreads and no dedicated issuer thread in the steady state: after an epilogue the tile's threads meet on a
// named barrier and ONE elected lane of one of their warps issues the next layer's tcgen05.mma's from warp-uniform code
// and commits them to the tile's mbarrier.  (A single issuer thread serving all tiles was measured at ~1200 cycles per
// layer step -- tools/field_trace.py -- because ptxas wrapped every UTCHMMA in a waterfall loop.)
//
//   forward   4 warpgroups x 1 tile, activations chained through TENSOR MEMORY (A operand in TMEM): see the kernel.
//   backward  2 slots x 1 tile (the activation tiles of a tile take 92 KB of shared memory), 256 threads per slot: two
//             threads per row, each owning 32 of a layer's 64 columns.  Per layer dgrad is issued FIRST and committed
//             separately, the weight-gradient MMAs by a second warp: the epilogue starts on the dgrad result while
//             they still run and only its in-place stores wait for them.
//
// Backward (autograd of the forward): recomputes the tile's forward with the forward kernel's arithmetic (identical
// activations, hence the ReLU masks the forward used), then walks the layers in reverse: wgrad = G^T X (contraction over
// the tile's 128 points, both operands viewed MN-major, M = 64, accumulators PERSISTENT in TMEM across every tile the
// CTA processes) and dgrad; the epilogue masks with [act > 0] and overwrites the dead activation tile with the gradient
// tile.  Hidden tiles carry a constant ones column, so the bias gradients come out of the wgrad MMAs themselves.
// Gradient tiles are fp16 carrying one global power-of-two scale (max|dL/dy| * 2^k in [2^9, 2^10)) measured by a
// max-reduction over dL/dy before the launch -- the job GradScaler does for the reference -- and divided out of d_enc
// and the weight gradients in fp32.  After its last tile the CTA adds its TMEM-resident weight gradients to global
// memory (one atomicAdd per weight per CTA).
//
// Operand-layout facts were established on hardware with tools/umma_probe.py (profiles/): for a canonical tile with
// row-group stride RG the K-major view is (lbo=128, sbo=RG, +256 B per K step), the MN-major view is (lbo=RG, sbo=128,
// +2*RG per K step); an MN-major A view wider than its tile reads on into the following shared memory and leaves
// garbage in the TMEM lanes beyond the tile's columns, which are never read; an M = 64 accumulator keeps row m in lane
// (m >> 4) * 32 + (m & 15); a TMEM A operand holds row r in lane r, two consecutive K elements per 32-bit column.
#include "field_mma.cuh"

namespace {

template <int E, bool TRACE>
__global__ void __launch_bounds__(FWG * 128, 1) k_field_fwd_mma(
    const __half* __restrict__ enc, const float* __restrict__ dirs, int dstride, int dgroup, int64_t P, int G,
    acn_field_weights w, float4* __restrict__ rgb_sigma, long long* __restrict__ trace, const int32_t* __restrict__ range)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    if (range) {            // rows [range[0], range[1]) only (an expert's bucket): P was just the launch's upper bound
        const int64_t r0 = __ldg(range);
        P = __ldg(range + 1) - r0;
        enc += r0 * E; dirs += r0 * dstride; rgb_sigma += r0;
    }
    if (P <= 0) return;     // an empty bucket: leave before staging weights / allocating TMEM
    using M = FwdMap<E>;
    constexpr WMap wm = wmap(E);
    constexpr int EC = E / 8;
    const uint32_t sb = umma::smem_u32(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5;

    stage_weights<E>(w, G, sb + M::w);
    if (tid == 0) {
        for (int i = 0; i < FWG; ++i) umma::mbar_init_a(sb + M::bars + 8 * i, 1);
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(reinterpret_cast<uint32_t*>(smem_raw + M::tmem_ptr), FWD_TMEM_COLS);
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_raw + M::tmem_ptr);

    const int wg = warp >> 2, row = tid & (TM - 1);
    const bool issuer_warp = (warp & 3) == wg;        // one issuing warp per warpgroup, each on its own SM sub-partition
    const uint32_t bar_id = 1 + wg;
    const uint32_t done = sb + M::bars + 8u * wg;
    const uint32_t H0 = tmem_base + (uint32_t)wg * 128u, H1 = H0 + 64u;     // issuer's view (lane 0)
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t h0 = H0 + lane_off, h1 = H1 + lane_off;                  // this thread's lanes
    const Tile Wt0 = mk_tile(sb + wm.t0, E), Wt1 = mk_tile(sb + wm.t1, 64), Whd = mk_tile(sb + wm.hd, 64),
               Wc0 = mk_tile(sb + wm.c0, 32), Wc1 = mk_tile(sb + wm.c1, 64), Wc2 = mk_tile(sb + wm.c2, 64);
    const Tile Bt0 = mk_tile(sb + wm.bt_t0, 16), Bt1 = mk_tile(sb + wm.bt_t1, 16), Bc0 = mk_tile(sb + wm.bt_c0, 16),
               Bc1 = mk_tile(sb + wm.bt_c1, 16), One = mk_tile(sb + wm.one, 16);

    const int64_t ntiles = (P + TM - 1) / TM;
    const int64_t tile_stride = (int64_t)gridDim.x * FWG;
    Tracer<TRACE> tr{ (trace && blockIdx.x == 0 && tid == 0) ? trace : nullptr, 0, TRACE_CAP };

    // every thread: my TMEM writes / reads are ordered before what the issuer launches next; then meet
    auto sync_group = [&]() { umma::fence_before_sync(); umma::bar_sync(bar_id, 128); };
    // one elected lane of the warpgroup's issuing warp launches a layer and commits it to `done`
    auto issue = [&](int step) {
        if (issuer_warp) {
            if (umma::elect_one()) {
                umma::fence_after_sync();
                switch (step) {
                    case 0: mma_layer_ts(H0, H1, Wt0, 64, E, &One, &Bt0); break;
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
    auto load_inputs = [&](int64_t tile, uint4* q, float* dir) {
        const int64_t p = tile * TM + row;
        const bool on = tile < ntiles && p < P;
        const int64_t ps = on ? p : 0;                     // a safe address; no select sits between the loads and their use
#pragma unroll
        for (int c = 0; c < EC; ++c) q[c] = __ldg(reinterpret_cast<const uint4*>(enc + ps * E) + c);
        const float* dp = dir_of(dirs, dstride, dgroup, ps);
        dir[0] = __ldg(dp); dir[1] = __ldg(dp + 1); dir[2] = __ldg(dp + 2);
    };

    uint32_t phase = 0;
    uint4 q[EC];
    float dir[3];
    int64_t tile = (int64_t)blockIdx.x * FWG + wg;
    load_inputs(tile, q, dir);
    float b_hd[16], b_c2[3];
#pragma unroll
    for (int qd = 0; qd < 4; ++qd) {
        const uint4 u = lds128(sb + wm.b_hd + 16 * qd);
        b_hd[4 * qd] = __uint_as_float(u.x); b_hd[4 * qd + 1] = __uint_as_float(u.y); b_hd[4 * qd + 2] = __uint_as_float(u.z); b_hd[4 * qd + 3] = __uint_as_float(u.w);
    }
    load_b_c2(sb + wm.b_c2, b_c2);

    for (; tile < ntiles; tile += tile_stride) {
        const int64_t p = tile * TM + row;
        tr(1);
        // ---- encoding row -> tensor memory (H1), layer 1 ----
#pragma unroll
        for (int c = 0; c < EC; c += 2) {
            const uint32_t r[8] = { q[c].x, q[c].y, q[c].z, q[c].w, q[c + 1].x, q[c + 1].y, q[c + 1].z, q[c + 1].w };
            umma::st8(h1 + c * 4, r);
        }
        umma::wait_st();
        const float cdir[3] = { dir[0], dir[1], dir[2] };
        sync_group(); issue(0);
        tr(2);
        load_inputs(tile + tile_stride, q, dir);       // prefetch: lands while this tile's layers run
        wait_done(done, phase); tr(3); epi_hidden_tmem(h0); tr(4); sync_group(); tr(5); issue(1); tr(6);     // -> trunk layer 2
        wait_done(done, phase); tr(7); epi_hidden_tmem(h1); sync_group(); issue(2);                          // -> heads
        wait_done(done, phase); tr(8);
        float sigma;
        {   // heads: accumulator columns 0..G-1 = geo, 15 = raw sigma; colour input row [sh(16) | geo(G) | 0] -> H0[0,16)
            float v[16], sh[16];
            umma::ld16(h0, v);
            sh16_fast(cdir[0], cdir[1], cdir[2], sh);
            umma::wait_ld();
            sigma = trunc_exp_fast(v[15] + b_hd[15]);
            uint32_t r[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = pack_h2(sh[2 * j], sh[2 * j + 1]);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = j < G ? v[j] + b_hd[j] : 0.0f;
#pragma unroll
            for (int j = 0; j < 8; ++j) r[8 + j] = pack_h2(v[2 * j], v[2 * j + 1]);
            umma::st16(h0, r);
            umma::wait_st();
        }
        sync_group(); issue(3);                                                                               // -> colour layer 1
        wait_done(done, phase); tr(9); epi_hidden_tmem(h1); sync_group(); issue(4);                          // -> colour layer 2
        wait_done(done, phase); tr(10); epi_hidden_tmem(h0); sync_group(); issue(5);                         // -> colour out
        wait_done(done, phase); tr(11);
        {
            float v[16];
            umma::ld16(h1, v);
            umma::wait_ld();
            if (p < P) rgb_sigma[p] = make_float4(sigmoid_fast(v[0] + b_c2[0]), sigmoid_fast(v[1] + b_c2[1]), sigmoid_fast(v[2] + b_c2[2]), sigma);
        }
        tr(12);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem_base, FWD_TMEM_COLS);
}

template <int E, bool TRACE, bool SCAT>
__global__ void __launch_bounds__(BNS * 256, 1) k_field_bwd_mma(
    const __half* __restrict__ enc, const float* __restrict__ dirs, int dstride, int dgroup, int64_t P, int G,
    acn_field_weights w, const float4* __restrict__ d_rgb_sigma, const unsigned int* __restrict__ absmax_bits,
    acn_field_grads g, float* __restrict__ d_enc, long long* __restrict__ trace, ScatterArgs sc, const int32_t* __restrict__ range)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    if (range) {            // rows [range[0], range[1]) only (an expert's bucket): P was just the launch's upper bound
        const int64_t r0 = __ldg(range);
        P = __ldg(range + 1) - r0;
        enc += r0 * E; dirs += r0 * dstride; d_rgb_sigma += r0;
        if (d_enc) d_enc += r0 * E;
        if (sc.x) sc.x += r0 * sc.xs;
    }
    if (P <= 0) return;     // an empty bucket: leave before staging weights / allocating TMEM
    using M = BwdMap<E>;
    using SM = SlotMap<E>;
    constexpr WMap wm = wmap(E);
    constexpr int EC = E / 8;            // 16-byte chunks per encoding row
    constexpr int ECH = (EC + 1) / 2;    // ... handled by one of the row's two threads
    const uint32_t sb = umma::smem_u32(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    stage_weights<E>(w, G, sb + M::w);
    for (int i = tid; i < BNS * 4 * TM; i += blockDim.x) {      // the ones chunk (chunk 8) of every hidden tile
        const int r = i & (TM - 1), t = (i >> 7) & 3, sl = i >> 9;
        const Tile ht = mk_tile(sb + M::slots + (uint32_t)sl * SM::bytes + SM::c2 + (uint32_t)t * (TM * HW * 2), HW);
        sts128(chunk_addr(ht, r, 8), make_uint4(0x00003C00u, 0u, 0u, 0u));
    }
    if (tid == 0) {
        for (int i = 0; i < BNS * 2; ++i) umma::mbar_init_a(sb + M::bars + 8 * i, 1);
        umma::fence_mbar_init();
    }
    if constexpr (SCAT) { if (tid < 16) umma::sts_f32(sb + M::res + 4u * tid, tid < sc.L ? (float)__ldg(sc.res + tid) : 1.0f); }
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

    const int slot = warp >> 3;                    // 8 warps per slot
    const int hcol = (warp >> 2) & 1;              // which half of a row's columns this thread owns
    const int row = (warp & 3) * 32 + lane;
    const bool issuer_warp = (warp & 7) == slot;      // slot 0 -> warp 0, slot 1 -> warp 9: different SM sub-partitions
    const uint32_t bar_id = 1 + slot;
    const uint32_t sbase = sb + M::slots + (uint32_t)slot * SM::bytes;
    const Tile Txe = mk_tile(sbase + SM::xe, E), Th1 = mk_tile(sbase + SM::h1, HW), Th2 = mk_tile(sbase + SM::h2, HW),
               Tcin = mk_tile(sbase + SM::cin, 32), Tc1 = mk_tile(sbase + SM::c1, HW), Tc2 = mk_tile(sbase + SM::c2, HW),
               Tdrr = mk_tile(sbase + SM::dg, 16), Tghd = mk_tile(sbase + SM::dg, 16);
    const Tile Wt0 = mk_tile(sb + M::w + wm.t0, E), Wt1 = mk_tile(sb + M::w + wm.t1, 64), Whd = mk_tile(sb + M::w + wm.hd, 64),
               Wc0 = mk_tile(sb + M::w + wm.c0, 32), Wc1 = mk_tile(sb + M::w + wm.c1, 64), Wc2 = mk_tile(sb + M::w + wm.c2, 64);
    const uint32_t wb = sb + M::w;
    const Tile Bt0 = mk_tile(wb + wm.bt_t0, 16), Bt1 = mk_tile(wb + wm.bt_t1, 16), Bc0 = mk_tile(wb + wm.bt_c0, 16),
               Bc1 = mk_tile(wb + wm.bt_c1, 16), One = mk_tile(wb + wm.one, 16);
    const uint32_t done_d = sb + M::bars + 16u * slot, done_w = done_d + 8u;
    const uint32_t dwin = tmem_base + (uint32_t)slot * 64u;
    const uint32_t tmem_d = dwin + ((uint32_t)((warp & 3) * 32) << 16);
    const int col0 = hcol * 32;

    const int64_t ntiles = (P + TM - 1) / TM;
    const int64_t tile_stride = (int64_t)gridDim.x * BNS;
    const float scale = grad_scale_from_max(__uint_as_float(__ldg(absmax_bits)));
    const float inv_scale = 1.0f / scale;
    const bool want_denc = SCAT || d_enc != nullptr;

    // Two warps of the slot issue: the chain issuer launches the MMAs the next epilogue waits for (forward layers, dgrad)
    // and commits them to done_d; a second warp, on another SM sub-partition, launches the weight-gradient MMAs of the
    // same step and commits them to done_w.  The two groups touch different accumulators, so their order is free; one
    // thread issuing all ~20 MMAs of a backward step made its warp ~800 cycles late to the slot's next barrier.
    const bool wgrad_warp = (warp & 7) == ((slot + 2) & 7);
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
                    default: if (want_denc) mma_dgrad(dwin, Th1, Wt0, E, 64); break;   // trunk layer 1 (optional d enc)
                }
                umma::commit_a(done_d);
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
    // this thread's share of a tile's inputs: its encoding chunks (c = hcol, hcol+2, ..), dir (hcol 1), dy (hcol 0)
    auto load_inputs = [&](int64_t tile, uint4* q, float* dir, float4& dy) {
        const int64_t p = tile * TM + row;
        const bool on = tile < ntiles && p < P;
#pragma unroll
        for (int j = 0; j < ECH; ++j) {
            const int c = 2 * j + hcol;
            q[j] = (on && c < EC) ? __ldg(reinterpret_cast<const uint4*>(enc + p * E) + c) : make_uint4(0u, 0u, 0u, 0u);
        }
        dir[0] = 0.f; dir[1] = 0.f; dir[2] = 1.f;
        dy = make_float4(0.f, 0.f, 0.f, 0.f);
        if (on) {
            if (hcol) { const float* dp = dir_of(dirs, dstride, dgroup, p); dir[0] = __ldg(dp); dir[1] = __ldg(dp + 1); dir[2] = __ldg(dp + 2); }
            else dy = __ldg(d_rgb_sigma + p);
        }
    };

    uint32_t ph_d = 0, ph_w = 0;
    uint4 encq[ECH];
    float dir[3];
    float4 dy;
    int64_t tile = (int64_t)blockIdx.x * BNS + slot;
    load_inputs(tile, encq, dir, dy);
    Tracer<TRACE> tr{ (trace && blockIdx.x == 0 && tid == 0) ? trace : nullptr, 0, TRACE_CAP };

    // Fused scatter, software-pipelined: the 16 encoding-gradient columns (8 levels) this thread holds at the end of a
    // tile are kept in registers and go out as REDs one level at a time in the first eight mbarrier waits of the NEXT tile
    // -- the atomics then drain through the LSU while the tensor core runs that tile's MMAs, instead of as one burst of
    // ~48 REDs per thread that stalls both slots' epilogues.  Measured on 2^24 samples: MLP backward + scatter as two
    // kernels 4.5 + 8.9 ms; burst at the end of the tile 11.2 ms; this 10.8 ms; + dedup of the coarse levels 10.0 ms.
    // Two ways to decouple the REDs from the chain further were tried and are SLOWER (patches under profiles/notes/):
    // four dedicated scatter warps reading d_enc from a TMEM window (21 ms: one warp per sub-partition is bound by the
    // latency of its ~5000 instructions per tile pair) and polling the chain's mbarriers with one scatter unit per
    // miss (11.9 ms: the out-of-line unit and 128 registers cost more than the idle slots give).  ncu of this version:
    // 1.08 G REDs, LSU wavefronts 51 %, L2 tags 46 %, tensor pipe 10 %, issue slots 33 % -- the kernel is bound by the
    // issue of the ~33 k warp instructions per tile pair with 16 resident warps, not by a memory unit.
    float pv[16];
    float ppos[3] = { 0.f, 0.f, 0.f };
    bool pon = false;
    const int lv0 = (16 * hcol) / 2;                      // first level of this thread's column group
    const bool has_cols = 16 * hcol < E;
    const uint32_t hmask = SCAT ? ((1u << sc.log2T) - 1u) : 0u;
    // Coarse levels (cells larger than the sample spacing: a ray stays in one cell for several samples, and the lanes of
    // a warp are 32 consecutive samples): the lanes of a run of equal cells add their 8 corner contributions with a
    // segmented warp scan and only the run's last lane issues the REDs -- fewer atomics, and no same-address collisions
    // inside one RED instruction (which the L2 atomic unit serialises).  The sum is over contiguous runs only, so it is
    // exact whatever the point order; at finer levels every lane scatters its own cell.
    constexpr int DEDUP_LEVELS = 4;
    auto drain = [&](int lv) {
        if constexpr (SCAT) {
            const int l = lv0 + lv;
            const float gx = pv[2 * lv], gy = pv[2 * lv + 1];
            const bool act = pon && !(gx == 0.0f && gy == 0.0f);      // fully occluded samples scatter nothing
            if (l < DEDUP_LEVELS) {                                    // warp-uniform (hcol is per warp)
                const GridCell c = grid_cell(ppos[0], ppos[1], ppos[2], umma::lds_f32(sb + M::res + 4u * l), sc.interp);
                const float wx[2] = { 1.0f - c.wx, c.wx }, wy[2] = { 1.0f - c.wy, c.wy }, wz[2] = { 1.0f - c.wz, c.wz };
                float2 acc[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float wk = act ? wz[k & 1] * wy[(k >> 1) & 1] * wx[(k >> 2) & 1] : 0.0f;
                    acc[k] = make_float2(gx * wk, gy * wk);
                }
                // cells of the coarse levels have < 2^10 cells per axis; an idle lane is a run of its own
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
                if (act && (lane == 31 || hnext))
                    scatter_cell_f2(reinterpret_cast<float2*>(sc.dtable) + ((size_t)l << sc.log2T), c.x0, c.y0, c.z0, hmask, acc);
                return;
            }
            if (!act) return;
            const GridCell c = grid_cell(ppos[0], ppos[1], ppos[2], umma::lds_f32(sb + M::res + 4u * l), sc.interp);
            const float wx[2] = { 1.0f - c.wx, c.wx }, wy[2] = { 1.0f - c.wy, c.wy }, wz[2] = { 1.0f - c.wz, c.wz };
            float2 acc[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float wk = wz[k & 1] * wy[(k >> 1) & 1] * wx[(k >> 2) & 1];
                acc[k] = make_float2(gx * wk, gy * wk);
            }
            scatter_cell_f2(reinterpret_cast<float2*>(sc.dtable) + ((size_t)l << sc.log2T), c.x0, c.y0, c.z0, hmask, acc);
        }
    };

    for (; tile < ntiles; tile += tile_stride) {
        const int64_t p = tile * TM + row;
        const bool on = p < P;
        // ---------------- forward recompute (same arithmetic as k_field_fwd_mma) ----------------
#pragma unroll
        for (int j = 0; j < ECH; ++j) if (2 * j + hcol < EC) sts128(chunk_addr(Txe, row, 2 * j + hcol), encq[j]);
        tr(1);
        group_sync(bar_id, 256); tr(2); issue(0); tr(3);
        const float cdir[3] = { dir[0], dir[1], dir[2] };
        const float4 cdy = dy;
        load_inputs(tile + tile_stride, encq, dir, dy);        // prefetch: lands while this tile runs
        drain(0);
        wait_done(done_d, ph_d); tr(4); epi_hidden32(tmem_d, col0, Th1, row); tr(5); group_sync(bar_id, 256); tr(6); issue(1); tr(7);
        drain(1);
        wait_done(done_d, ph_d); epi_hidden32(tmem_d, col0, Th2, row); group_sync(bar_id, 256); issue(2);
        drain(2);
        wait_done(done_d, ph_d);
        float sig_raw = 0.0f;
        if (hcol) epi_heads_sh(cdir, Tcin, row);
        else sig_raw = epi_heads_geo(tmem_d, wb + wm.b_hd, G, Tcin, row, 1.0f);
        group_sync(bar_id, 256); issue(3);
        drain(3);
        wait_done(done_d, ph_d); epi_hidden32(tmem_d, col0, Tc1, row); group_sync(bar_id, 256); issue(4);
        drain(4);
        wait_done(done_d, ph_d); epi_hidden32(tmem_d, col0, Tc2, row); group_sync(bar_id, 256); issue(5);
        drain(5);
        wait_done(done_d, ph_d);
        float d_sig = 0.0f;
        if (!hcol) {   // output gradients (scaled): d rgb_raw = dy * y (1 - y); d sigma_raw = dy * exp(clamp(sigma_raw))
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
        tr(8);
        group_sync(bar_id, 256); tr(9); issue(6); tr(10);
        drain(6);
        // ---------------- backward ----------------
        uint4 o[4];
        wait_done(done_d, ph_d); tr(11); epi_mask32_load(tmem_d, col0, Tc2, row, o); tr(12); wait_done(done_w, ph_w); tr(13); epi_mask32_store(Tc2, row, col0, o);
        tr(14); group_sync(bar_id, 256); tr(15); issue(7); tr(16);
        drain(7);
        wait_done(done_d, ph_d); epi_mask32_load(tmem_d, col0, Tc1, row, o); wait_done(done_w, ph_w); epi_mask32_store(Tc1, row, col0, o);
        group_sync(bar_id, 256); issue(8);
        wait_done(done_d, ph_d);
        if (!hcol) {   // d cin = [d sh | d geo | 0] -> heads gradient row [d geo | 0 | d sigma_raw @15]
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
        group_sync(bar_id, 256); issue(9);
        wait_done(done_d, ph_d); epi_mask32_load(tmem_d, col0, Th2, row, o); wait_done(done_w, ph_w); epi_mask32_store(Th2, row, col0, o);
        group_sync(bar_id, 256); issue(10);
        float upos[3] = { 0.f, 0.f, 0.f };
        if constexpr (SCAT) {   // the point's unit-cube position for the table scatter: in flight while the last two layers run
            if (on) {
                if (sc.rays) {
                    const int64_t ray = p / sc.S;
                    const float* ry = sc.rays + 8 * ray;
                    const float tv = __ldg(sc.t + p);
#pragma unroll
                    for (int c = 0; c < 3; ++c) upos[c] = __fadd_rn(__ldg(ry + c), __fmul_rn(__ldg(ry + 3 + c), tv));
                } else {
#pragma unroll
                    for (int c = 0; c < 3; ++c) upos[c] = __ldg(sc.x + p * sc.xs + c);
                }
                if (sc.box6) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) upos[c] = world_to_unit1(upos[c], __ldg(sc.box6 + c), __ldg(sc.box6 + 3 + c));
                }
            }
        }
        wait_done(done_d, ph_d); epi_mask32_load(tmem_d, col0, Th1, row, o); wait_done(done_w, ph_w); epi_mask32_store(Th1, row, col0, o);
        group_sync(bar_id, 256); issue(11);
        wait_done(done_d, ph_d);
        if constexpr (SCAT) {   // d_enc never leaves the SM: this thread's 8 levels are stashed for drain() during the next tile
            if (has_cols) {
                float v[16];
                umma::ld16(tmem_d + 16 * hcol, v);
                umma::wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) pv[j] = v[j] * inv_scale;
                ppos[0] = upos[0]; ppos[1] = upos[1]; ppos[2] = upos[2];
                pon = on;
            }
        } else if (want_denc) {   // 16-column groups of d_enc alternate between the row's two threads
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int c0 = 16 * (2 * j + hcol);
                if (c0 < E) {
                    float v[16];
                    umma::ld16(tmem_d + c0, v);
                    umma::wait_ld();
                    if (on) {
                        float4* dst = reinterpret_cast<float4*>(d_enc + p * E + c0);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            dst[k] = make_float4(v[4 * k] * inv_scale, v[4 * k + 1] * inv_scale, v[4 * k + 2] * inv_scale, v[4 * k + 3] * inv_scale);
                    }
                }
            }
        }
        wait_done(done_w, ph_w);     // the xe / h1 tiles are free again only now
        tr(17);
    }

    if constexpr (SCAT) {
#pragma unroll
        for (int lv = 0; lv < 8; ++lv) drain(lv);
    }
    // ---------------- add the TMEM-resident weight gradients to global memory ----------------
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    if (warp < 4) {
        const uint32_t tmem_row = tmem_base + ((uint32_t)(warp * 32) << 16);
        // M = 64 accumulators: rows 16w..16w+15 live in lanes 32w..32w+15 (the upper half of each warp holds nothing);
        // M = 128 accumulators: row = lane.
        const int t = (lane < 16) ? warp * 16 + lane : 1 << 20;     // row of this lane in an M=64 accumulator (or none)
        const int u = tid;                                           // row of this lane in an M=128 accumulator
        // accumulator(this lane, column c0 + j) -> dst[j * ld_col] for j < nvalid (dst null: skip); ncols multiple of 16
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
        // [dW | db] (out x in+1) accumulators, M = 64: lane = output unit, columns 0..63 = input units, column 64 = bias
        flush(COL_WC1, 64, at(g.p[10], t < 64, t * 64), 1, 64);
        flush(COL_WC1 + 64, 16, at(g.p[11], t < 64, t), 1, 1);
        flush(COL_WT1, 64, at(g.p[2], t < 64, t * 64), 1, 64);
        flush(COL_WT1 + 64, 16, at(g.p[3], t < 64, t), 1, 1);
        flush(COL_WT0, 64, at(g.p[0], t < 64, t * E), 1, E);
        flush(COL_BT0, 16, at(g.p[1], t < 64, t), 1, 1);
        // colour layer 1: tile column k' is source column cin_src_col(k'): SH block, geo block, then the ones column (31)
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
        // [dW^T ; db] (in+1 x out) accumulators, M = 128: lane = input unit (lane 64 = bias), column = output unit
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
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem_base, BWD_TMEM_COLS);
}

#ifdef ACN_DEBUG_BUILD
long long* g_field_trace = nullptr;   // libacn_b200_debug.so only: set by acn_debug_field_trace for the following launches
constexpr bool kTraceBuild = true;
#else
constexpr long long* g_field_trace = nullptr;   // the product library has no timeline build of the kernels
constexpr bool kTraceBuild = false;
#endif

template <int E>
int launch_fwd(acn_ctx* ctx, const void* enc, const float* dirs, int dirs_stride, int dirs_group, int64_t P, int G,
               const acn_field_weights* w, float* rgb_sigma, const int32_t* range, cudaStream_t st) {
    constexpr uint32_t smem = FwdMap<E>::bytes;
    ACN_REQUIRE((int)smem <= ctx->max_smem_optin, ACN_EUNSUPPORTED, "acn_field_fwd(f16): needs %u B shared memory", smem);
    const int64_t ntiles = (P + TM - 1) / TM;
    int64_t grid = (ntiles + FWG - 1) / FWG;
    if (grid > ctx->sm_count) grid = ctx->sm_count;
    if (kTraceBuild && g_field_trace && E == 32) {          // the timeline build of the kernel (tools/field_trace.py)
        ACN_CUDA(cudaFuncSetAttribute(k_field_fwd_mma<32, kTraceBuild>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_field_fwd_mma<32, kTraceBuild><<<(int)grid, FWG * 128, smem, st>>>((const __half*)enc, dirs, dirs_stride, dirs_group, P, G, *w,
                                                                       (float4*)rgb_sigma, g_field_trace, range);
    } else {
        ACN_CUDA(cudaFuncSetAttribute(k_field_fwd_mma<E, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_field_fwd_mma<E, false><<<(int)grid, FWG * 128, smem, st>>>((const __half*)enc, dirs, dirs_stride, dirs_group, P, G, *w,
                                                                       (float4*)rgb_sigma, nullptr, range);
    }
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

template <int E>
int launch_bwd(acn_ctx* ctx, const void* enc, const float* dirs, int dirs_stride, int dirs_group, int64_t P, int G,
               const acn_field_weights* w, const float* d_rgb_sigma, const unsigned int* absmax, const acn_field_grads* g,
               float* d_enc, const ScatterArgs* sc, const int32_t* range, cudaStream_t st) {
    constexpr uint32_t smem = BwdMap<E>::bytes;
    ACN_REQUIRE((int)smem <= ctx->max_smem_optin, ACN_EUNSUPPORTED, "acn_field_bwd(f16): needs %u B shared memory", smem);
    const int64_t ntiles = (P + TM - 1) / TM;
    int64_t grid = (ntiles + BNS - 1) / BNS;
    if (grid > ctx->sm_count) grid = ctx->sm_count;
    const ScatterArgs none{};
#ifdef ACN_DEBUG_BUILD
    if (sc) {
        if constexpr (E <= 32) {             // the single-role fused table scatter (debug library): E = 2 L, L <= 16
            ACN_CUDA(cudaFuncSetAttribute(k_field_bwd_mma<E, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_field_bwd_mma<E, false, true><<<(int)grid, BNS * 256, smem, st>>>((const __half*)enc, dirs, dirs_stride, dirs_group, P, G, *w,
                                                                                 (const float4*)d_rgb_sigma, absmax, *g, nullptr, nullptr, *sc, range);
        } else {
            ACN_REQUIRE(false, ACN_EUNSUPPORTED, "acn_render_expert_bwd: encoding width %d > 32", E);
        }
    } else
#endif
    if (kTraceBuild && g_field_trace && E == 32) {          // the timeline build of the kernel (tools/field_trace.py --bwd)
        ACN_CUDA(cudaFuncSetAttribute(k_field_bwd_mma<32, kTraceBuild, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_field_bwd_mma<32, kTraceBuild, false><<<(int)grid, BNS * 256, smem, st>>>((const __half*)enc, dirs, dirs_stride, dirs_group, P, G, *w,
                                                                              (const float4*)d_rgb_sigma, absmax, *g, d_enc, g_field_trace, none, range);
    } else {
        ACN_CUDA(cudaFuncSetAttribute(k_field_bwd_mma<E, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_field_bwd_mma<E, false, false><<<(int)grid, BNS * 256, smem, st>>>((const __half*)enc, dirs, dirs_stride, dirs_group, P, G, *w,
                                                                              (const float4*)d_rgb_sigma, absmax, *g, d_enc, nullptr, none, range);
    }
    ACN_CHECK_LAUNCH();
    return ACN_OK;
}

}  // namespace

#ifdef ACN_DEBUG_BUILD
extern "C" int acn_debug_field_trace(acn_ctx* ctx, long long* trace_or_null) {
    ACN_CHECK_CTX(ctx);
    g_field_trace = trace_or_null;
    return ACN_OK;
}
#endif

int acn_field_fwd_tc(acn_ctx* ctx, const void* enc, int enc_dtype, const float* dirs, int dirs_stride, int dirs_group,
                     int64_t P, int E, int H, int G, int C, const acn_field_weights* w, float* rgb_sigma, const int32_t* range,
                     cudaStream_t st) {
    int rc = check_dims("acn_field_fwd", enc_dtype, E, H, G, C, enc);
    if (rc) return rc;
    switch (E) {
        case 16: return launch_fwd<16>(ctx, enc, dirs, dirs_stride, dirs_group, P, G, w, rgb_sigma, range, st);
        case 32: return launch_fwd<32>(ctx, enc, dirs, dirs_stride, dirs_group, P, G, w, rgb_sigma, range, st);
        case 48: return launch_fwd<48>(ctx, enc, dirs, dirs_stride, dirs_group, P, G, w, rgb_sigma, range, st);
        default: return launch_fwd<64>(ctx, enc, dirs, dirs_stride, dirs_group, P, G, w, rgb_sigma, range, st);
    }
}

int acn_field_bwd_tc(acn_ctx* ctx, const void* enc, int enc_dtype, const float* dirs, int dirs_stride, int dirs_group,
                     int64_t P, int E, int H, int G, int C, const acn_field_weights* w, const float* d_rgb_sigma,
                     const acn_field_grads* g, float* d_enc, const int32_t* range, cudaStream_t st) {
    int rc = check_dims("acn_field_bwd", enc_dtype, E, H, G, C, enc);
    if (rc) return rc;
    ACN_REQUIRE(((uintptr_t)d_enc & 15) == 0, ACN_EINVAL, "acn_field_bwd(f16): d_enc must be 16-byte aligned");
    unsigned int* slot = nullptr;
    rc = absmax_word(ctx, "acn_field_bwd", d_rgb_sigma, P, range, st, &slot);
    if (rc) return rc;
    switch (E) {
        case 16: return launch_bwd<16>(ctx, enc, dirs, dirs_stride, dirs_group, P, G, w, d_rgb_sigma, slot, g, d_enc, nullptr, range, st);
        case 32: return launch_bwd<32>(ctx, enc, dirs, dirs_stride, dirs_group, P, G, w, d_rgb_sigma, slot, g, d_enc, nullptr, range, st);
        case 48: return launch_bwd<48>(ctx, enc, dirs, dirs_stride, dirs_group, P, G, w, d_rgb_sigma, slot, g, d_enc, nullptr, range, st);
        default: return launch_bwd<64>(ctx, enc, dirs, dirs_stride, dirs_group, P, G, w, d_rgb_sigma, slot, g, d_enc, nullptr, range, st);
    }
}

// The SINGLE-ROLE fused backward (k_field_bwd_mma<.., SCAT>: every MLP thread scatters its own d enc columns between its
// epilogues), superseded by the warp-specialised kernel of expert_bwd.cu (8.9 vs 10.0 ms).  Debug library only: the
// cross-check / A-B partner of acn_render_expert_bwd (tools/prof_fused_bwd.py, tests/test_gpu_tc.py); same arguments.
#ifdef ACN_DEBUG_BUILD
extern "C" int acn_debug_render_expert_bwd_single(acn_ctx* ctx, const float* x_or_null, int x_stride, const float* rays8_or_null,
                                     const float* t_vals_or_null, int64_t P, int S, const int32_t* range_or_null,
                                     const float* box6_or_null, int L, int F,
                                     int log2T, const int32_t* res, int interp, const void* enc_f16, const float* dirs,
                                     int dirs_stride, int dirs_group, int H, int G, int C, const acn_field_weights* w,
                                     const float* d_rgb_sigma, const acn_field_grads* g, float* dtable, acn_stream stream) {
    ACN_CHECK_CTX(ctx);
    const char* fn = "acn_debug_render_expert_bwd_single";
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
    rc = absmax_word(ctx, fn, d_rgb_sigma, P, range_or_null, st, &slot);
    if (rc) return rc;
    const ScatterArgs sc{ from_rays ? nullptr : x_or_null, x_stride, rays8_or_null, t_vals_or_null, S, box6_or_null, dtable, L, log2T, res, interp };
    if (E == 16) return launch_bwd<16>(ctx, enc_f16, dirs, dirs_stride, dirs_group, P, G, w, d_rgb_sigma, slot, g, nullptr, &sc, range_or_null, st);
    return launch_bwd<32>(ctx, enc_f16, dirs, dirs_stride, dirs_group, P, G, w, d_rgb_sigma, slot, g, nullptr, &sc, range_or_null, st);
}
#endif
