// Fused J/K kernel over the native v2 layout (layout.cuh).  Included by jk_fused.cu.
//
// A WARP is the unit of work.  A warp task is (16 sub-rows k) x (one 64-column chunk, two columns per lane) x (a list
// of row-blocks: 4 rows i x NB rows j of one tile, NB = 4 for the single-density passes, 2 for the dual-density UHF
// pass).  In the v2 layout the R = 4 NB rows of a row-block for one (k, chunk) are ONE contiguous piece and the 16
// sub-rows of a task follow each other, so the ERI stream of a task is a single running pointer:
//   * full pieces (64 columns): one elected lane issues ONE cp.async.bulk (TMA) per pipeline stage into the warp's
//     shared-memory ring, completion on a per-stage mbarrier (UBLKCP in SASS);
//   * the triangular edge (pieces of 16 / 32 / 48 columns, k - 64 c < 48): 16-byte cp.async per lane with immediate
//     row offsets and the zero-fill form for the lanes past the piece, landing in the same [R][64] stage layout.
// Cells outside the canonical range are stored as 0.0, so the consumer is mask free and has no per-row address
// arithmetic at all.
//
// Consumer, per sub-row: every lane holds D[j,l], D[i,l] (its two columns), the accumulators K~[i,l], K~[j,l], J~[i,j];
// D[i,k], D[j,k] and D2[i,j] are warp-uniform (staged per row-block in shared memory, broadcast loads); J~[k,l] goes
// to a warp-private slab that lives for the whole task.  K~[i,k], K~[j,k] need a sum over the lanes once per sub-row;
// that reduction is SOFTWARE-PIPELINED over three iterations so that the loop body is one straight-line block in which
// ptxas can place all the shuffles / shared-memory round trips in the issue slots the half-rate DFMAs leave free:
//     iteration kk:   A  the 12 R DFMAs of sub-row kk                                    -> lane partials cd(kk)
//                     B  xor-16 shuffle of cd(kk-1), 16 lanes store them (transposed)     -> s_red[(kk-1) & 1]
//                     C  partials of kk-2: 4 loads, 3 adds, 2 shuffles, one store         -> s_kout[.][kk-2]
// and the 16 x NV reduced values of a row-block leave with coalesced red.global.add.f64 after its sweep.
#pragma once

namespace sqcc {

__device__ __forceinline__ uint32_t smem_u32(const void *ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
// The ERI stream is read exactly once per build: it goes through L2 with the evict-first policy so that it does not push
// out the density / accumulator matrices every warp keeps coming back to (with the default policy a line lives ~20 us
// in the 126 MB L2 at 6.5 TB/s -- shorter than the time between two visits of a D2 tile row by the same warp).
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
template <bool HINT>
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void *src, uint32_t src_bytes, uint64_t pol)
{
    if (HINT)
        asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;" ::"r"(dst), "l"(src), "r"(src_bytes), "l"(pol)
                     : "memory");
    else
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "MBAR_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra MBAR_DONE_%=;\n\t"
        "bra MBAR_WAIT_%=;\n\t"
        "MBAR_DONE_%=:\n\t}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// one contiguous global -> shared bulk copy (TMA), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t pol)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "l"(pol)
                 : "memory");
}
// 16-byte read-only load that asks L2 to keep the line (density tiles)
__device__ __forceinline__ double2 ldg2_keep(const double *p, uint64_t pol)
{
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(r.x), "=d"(r.y) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
// xor-shuffle of a double without the volatile mov.b64 pack/unpack of the CUDA header
__device__ __forceinline__ double shfl_xor_f64(double v, int lane_mask)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xffffffffu, lo, lane_mask);
    hi = __shfl_xor_sync(0xffffffffu, hi, lane_mask);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Flushes to the global work matrices.  FIXED = false: red.global.add.f64 (result depends on the arrival order in the
// last bits).  FIXED = true: the value is converted to 64-bit fixed point (scale 2^fx_shift, chosen by the prep kernel
// from the operand magnitudes) and added with an integer red: integer addition is associative, so the sums -- and
// with them J and K -- are bit-for-bit reproducible from run to run and independent of the task schedule.
template <bool FIXED>
__device__ __forceinline__ void flush_add(bool pred, double *ptr, double v, double fx_scale)
{
    if (pred) {
        if (FIXED) {
            const long long q = __double2ll_rn(v * fx_scale);
            asm volatile("red.global.add.u64 [%0], %1;" ::"l"(ptr), "l"(q) : "memory");
        } else {
            asm volatile("red.global.add.f64 [%0], %1;" ::"l"(ptr), "d"(v) : "memory");
        }
    }
}

constexpr int JK_RED_STRIDE = 17;  // value stride of the reduction scratch (odd: conflict-free 4x4 read pattern)
constexpr int JK_RED_WIDE = 40;    // 32-entry rows: stride == 8 (mod 16) doubles, see JKSmem::WIDE_RED

// NH: ring slots per sub-row.  NH = 2 (4 x 4 row-blocks): a slot holds HALF a piece (8 rows, 4 KB) and is handed back to
// the producer as soon as its 8 rows are consumed, so the same 16 KB ring keeps up to 3 half-pieces (12 KB instead of 8 KB)
// in flight -- the per-sub-row time is (latency + compute) / slots until compute bounds it.
// SPIN2 (dual-density UHF pass): the two half-warps work on the SAME integrals (16 rows x 32 columns, two
// columns per lane) for the two spin densities, so every per-spin array exists once per lane (NS = 1) and the per-spin
// shared-memory arrays twice per warp.
// KS (HM = 2, narrow pieces of the triangular edge, widths 16 / 32): the two half-warps work on two CONSECUTIVE SUB-ROWS
// of the same 32 columns, so a loop iteration -- whose cost is its 12 x 16 DFMAs whatever the piece width -- covers two
// sub-rows instead of one.
template <int NB, int NS, int STAGES, bool SLAB, int NH, int HM>
struct JKSmem {
    static constexpr bool SPIN2 = HM == 1, KS = HM == 2, HALF = HM != 0;
    static constexpr int R = 4 * NB;
    static constexpr int HR = R / NH;                 // rows per ring slot
    static constexpr int NV = NS * (4 + NB);          // lane-partial values reduced per sub-row (per half-warp if SPIN2)
    static constexpr int NVE = (NV + 1) & ~1;
    static constexpr int NSP = SPIN2 ? 2 : 1;                // spin planes staged in s_u
    static constexpr int NRV = HALF ? 2 * NV : NV;           // values reduced per iteration (both half-warps)
    static constexpr int ROW_BYTES = HALF ? 256 : 512;       // one row of a ring slot / of the J~ slab
    static constexpr int STAGE_BYTES = (KS ? 2 : 1) * HR * ROW_BYTES;
    // WIDE_RED (4 x 2 dual-density pass, whole-warp chunks): the reduction scratch keeps all 32 lane partials of a value
    // (no xor-16 fold before the store), rows of JK_RED_WIDE doubles read with conflict-free 128-bit loads, ONE buffer (the
    // loads of an iteration are issued before the warp barrier its stores follow).
    static constexpr bool WIDE_RED = NB == 2 && HM == 0;
    static constexpr int RED_STRIDE = WIDE_RED ? JK_RED_WIDE : JK_RED_STRIDE;
    static constexpr int RED_NBUF = WIDE_RED ? 1 : 2;
    // doubles per buffer (>= 8 values: J~ row dots go 8 at a time; the last wide row needs no padding)
    static constexpr int RED_BUF = WIDE_RED ? (NRV - 1) * JK_RED_WIDE + 32 : (NRV > 8 ? NRV : 8) * RED_STRIDE;
    static constexpr int OFF_SLAB = 0;                                   // double2 [16][32 | 16]
    static constexpr int OFF_RING = OFF_SLAB + (SLAB ? JK_KN * ROW_BYTES : 0); // [STAGES][HR][ROW_BYTES]
    static constexpr int OFF_U = OFF_RING + STAGES * STAGE_BYTES;        // double [NSP][16][NVE]
    static constexpr int OFF_RED = OFF_U + NSP * JK_KN * NVE * 8;        // double [RED_NBUF][RED_BUF]
    static constexpr int OFF_KOUT = OFF_RED + RED_NBUF * RED_BUF * 8;    // double [NRV][17]
    static constexpr int OFF_W = OFF_KOUT + ((NRV + 1) & ~1) * JK_RED_STRIDE * 8;     // double [R]
    static constexpr int OFF_BAR = OFF_W + R * 8;                        // u64 [STAGES]
    // RBS (dual-density pass): the row-block list of the task lives in shared memory instead of one entry per lane +
    // shuffles -- the pass has no registers to spare, and a spilled register is an L2 round trip
    static constexpr bool RBS = NB == 2 && HM == 0;
    static constexpr int OFF_RB = (OFF_BAR + STAGES * 8 + 15) & ~15;     // RowBlock [32]
    // (same pass) D2S: the density tile rows D2[k0 + kk, chunk] come through a two-slot cp.async ring in shared memory, one
    // sub-row ahead: as register loads they shared a scoreboard with the first shared-memory loads of the next FMA block,
    // which then waited for the L2 round trip
    static constexpr int OFF_D2S = OFF_RB + (RBS ? 512 : 0);              // double2 [2][32]
    static constexpr int WARP_BYTES = (OFF_D2S + (RBS ? 1024 : 0) + 63) & ~63;
};

// NB: rows j per row-block (4: RHF / J-only, 2: UHF).  TMA: full pieces through cp.async.bulk (false: cp.async only).
// HINT: L2 cache-policy hints (evict-first on the ERI stream, evict-last on the density tile rows).
// SLAB: J~[k, l-chunk] accumulates over the row-blocks of a task in a warp-private shared-memory slab (true) or leaves
// with two red.global per lane per sub-row (false: the 8 KB pay for one more ring stage).
// SPIN2: dual-density pass with the spins on the two half-warps (NB = 4, cp.async ring, 32-column chunks: task.chunk counts
// 32-column chunks); lane = 16 spin + column pair.
// HM = 2 (KS): narrow-piece pass (tasks whose pieces are 16 or 32 columns wide), half-warp = sub-row parity.
template <int NB, int NSPIN, bool WANTK, int WARPS, int STAGES, bool TMA, bool FIXED, bool HINT, bool SLAB, int NH, int HM>
__global__ void __launch_bounds__(WARPS * 32, 1) jk_stream_kernel(const JKParams p)
{
    constexpr int NI = 4;
    constexpr bool SPIN2 = HM == 1, KS = HM == 2, HALF = HM != 0;
    constexpr int NS = (WANTK && !SPIN2) ? NSPIN : 1;
    using SM = JKSmem<NB, NS, STAGES, SLAB, NH, HM>;
    constexpr int R = SM::R, HR = SM::HR, NV = SM::NV, NVE = SM::NVE;
    static_assert(NB % NH == 0, "a ring slot holds whole j-columns of the row-block");
    static_assert(!SPIN2 || (NB == 4 && NH == 1 && !TMA && WANTK && NSPIN == 2), "spin-split pass: 4 x 4 row-blocks, cp.async ring");
    static_assert(!KS || (NH == 1 && !TMA), "narrow-piece pass: cp.async ring");
    constexpr int LW = HALF ? 16 : 32;                // double2 per ring-slot row = lanes that share one row
    constexpr int KSTEP = KS ? 2 : 1;                 // sub-rows per loop iteration
    constexpr int NVP = (NV + 7) / 8;
    constexpr int NHV = 2 * NV;   // half-warp modes: values reduced per iteration, 16 lane partials each (16, or 24 dual-density)
    static_assert(!HALF || NHV == 16 || NHV == 24, "half-warp reduction: one pass of 16 values, a second one of 8");
    constexpr bool DEEP = NB == 4;   // reduction pipeline over three iterations (needs NV more live registers) or two
    constexpr bool WIDE = SM::WIDE_RED;   // 32 lane partials per value in the scratch, no xor-16 fold (JKSmem)
    constexpr int RW = WIDE ? 8 : 4;      // scratch entries one lane sums per value
    static_assert(!WIDE || (!DEEP && !HALF), "wide reduction rows: shallow pipeline, whole-warp chunks");
    constexpr uint32_t STAGE_BYTES = SM::STAGE_BYTES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ql = HALF ? (lane & 15) : lane;         // column pair of this lane inside the chunk
    const int hh = HALF ? (lane >> 4) : 0;            // half-warp: spin density (SPIN2) / sub-row parity (KS)
    const int sp = SPIN2 ? hh : 0;                    // spin density this lane works for
    unsigned char *wbase = smem_raw + (size_t)warp * SM::WARP_BYTES;
    double2 *s_jt = reinterpret_cast<double2 *>(wbase + SM::OFF_SLAB);
    const double2 *ring = reinterpret_cast<const double2 *>(wbase + SM::OFF_RING);
    double *s_u = reinterpret_cast<double *>(wbase + SM::OFF_U);
    double *s_red = reinterpret_cast<double *>(wbase + SM::OFF_RED);
    double *s_kout = reinterpret_cast<double *>(wbase + SM::OFF_KOUT);
    double *s_w = reinterpret_cast<double *>(wbase + SM::OFF_W);
    RowBlock *s_rb = reinterpret_cast<RowBlock *>(wbase + SM::OFF_RB);
    constexpr bool RBS = SM::RBS, D2S = SM::RBS;
    double2 *s_d2 = reinterpret_cast<double2 *>(wbase + SM::OFF_D2S);
    uint32_t d2slot = 0;   // (D2S) ring slot the next iteration reads
    const uint32_t ring_u32 = smem_u32(wbase + SM::OFF_RING);
    const uint32_t bar_u32 = smem_u32(wbase + SM::OFF_BAR);
    const int n = p.n, ldd = p.ldd;
    const size_t nl = (size_t)n * ldd;
    const double fx = p.fx_scale;
    const double *dsp_l = p.dsp + (size_t)sp * nl;    // density / accumulator plane of this lane's spin
    double *kt_l = p.kt + (size_t)sp * nl;
    const uint64_t pol_stream = HINT ? l2_policy_evict_first() : 0, pol_keep = HINT ? l2_policy_evict_last() : 0;
    uint32_t stage_w = 0, stage_r = 0, phase_bits = 0;   // ring write / read positions and mbarrier parities, persist across tasks

    if (TMA) {
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < STAGES; ++s) mbar_init(bar_u32 + 8 * s, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
    }

    for (;;) {
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(p.counter, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= (unsigned)p.ntasks) break;
        const JKTask task = p.tasks[t];
        const int k0 = task.k0;
        const int c = SPIN2 ? task.chunk >> 1 : task.chunk;        // 64-column chunk of the layout
        const int xcol = (SPIN2 ? (task.chunk & 1) * 32 : 0) + 2 * ql;   // this lane's first column inside the piece
        const int cbase = c * JK_CHUNK;
        const int l0 = cbase + xcol;
        const bool lane_in = l0 < ldd;
        const int w = lay_w(k0 - cbase);                 // piece width of this task (uniform: k0, cbase multiples of 16)
        const bool full = TMA && w == JK_CHUNK;

        if (SLAB) {
#pragma unroll 4
            for (int kk = 0; kk < JK_KN; ++kk) s_jt[kk * LW + ql] = make_double2(0.0, 0.0);
        }

        // row-block list of the task, one entry per lane (rb_count <= 32).  The tasks that share a group of row-blocks
        // (other k-ranges / chunks) are neighbours in the queue and run at the same time on other warps; walking the list
        // in the same order they would all flush K~[j, l] / K~[j, k] of the SAME rows j at the same moment, and atomics on
        // one address serialise in L2.  Every task therefore starts at its own offset into the (cyclic) list.
        long long my_base = 0;
        int my_imax = k0, my_j0 = 0;
        if (lane < task.rb_count) {
            int src = lane + (p.rotate ? (int)((t * 40503u >> 3) % (unsigned)task.rb_count) : 0);
            if (src >= task.rb_count) src -= task.rb_count;
            const RowBlock rbk = p.rowblocks[task.rb_begin + src];
            my_base = rbk.base;
            my_imax = rbk.imax;
            my_j0 = rbk.j0;
        }
        if (RBS) {
            s_rb[lane] = RowBlock{my_base, my_imax, my_j0};
            __syncwarp();
        }
        auto rb_imax = [&](int idx) { return RBS ? s_rb[idx].imax : __shfl_sync(0xffffffffu, my_imax, idx); };
        auto rb_j0 = [&](int idx) { return RBS ? s_rb[idx].j0 : __shfl_sync(0xffffffffu, my_j0, idx); };
        auto rb_base = [&](int idx) { return RBS ? (long long)s_rb[idx].base : __shfl_sync(0xffffffffu, my_base, idx); };
        // sub-rows to stream in this task
        int total_it = 0;
        {
            int cnt = lane < task.rb_count ? (my_imax + 1 - k0 < JK_KN ? my_imax + 1 - k0 : JK_KN) : 0;
            if (KS) cnt = (cnt + 1) >> 1;   // two sub-rows per iteration
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            total_it = cnt;
        }

        if (D2S) cp_async_wait<0>();   // (the last look-ahead row of the previous task may still be on its way into the ring)
        // ---- producer: one piece (sub-row) per call
        int prod_rbi = -1, prod_left = 0, prod_todo = total_it * NH, prod_half = 0;   // prod_todo counts ring slots
        const char *prod_src = reinterpret_cast<const char *>(p.P);
        const uint32_t lane_bytes = xcol < w ? 16u : 0u;
        auto issue_next = [&]() {
            if (prod_left == 0) {   // next row-block of the list
                ++prod_rbi;
                const long long base = rb_base(prod_rbi);
                const int imax = rb_imax(prod_rbi);
                const int j0 = rb_j0(prod_rbi);
                prod_left = imax + 1 - k0 < JK_KN ? imax + 1 - k0 : JK_KN;
                long long off = base + lay_piece_offset(imax, c, k0);
                if (NB == 2) off += (long long)((j0 >> 1) & 1) * 8 * w;   // second 4 x 2 half of the tile's pieces
                prod_src = reinterpret_cast<const char *>(p.P + off);
            }
            const uint32_t dst = ring_u32 + stage_w * STAGE_BYTES;
            const char *half_src = prod_src + (size_t)prod_half * (HR * 8) * w;   // rows [prod_half * HR, +HR) of the piece
            if (full) {
                if (lane == 0) {
                    const uint32_t bar = bar_u32 + 8 * stage_w;
                    mbar_expect_tx(bar, STAGE_BYTES);
                    if (HINT)
                        bulk_g2s_hint(dst, half_src, STAGE_BYTES, bar, pol_stream);
                    else
                        bulk_g2s(dst, half_src, STAGE_BYTES, bar);
                }
            } else if (KS) {
                // two consecutive narrow pieces, one per half-warp: lane (hh, ql) brings its own 16 bytes of the 16 rows of
                // sub-row 2 it + hh (zero-filled past the piece width and past the last sub-row of the sweep)
                const uint32_t nb = (xcol < w && hh < prod_left) ? 16u : 0u;
                const char *src = half_src + 16 * ql + (size_t)(hh < prod_left ? hh : 0) * (LAY_ROWS * 8) * w;
                const uint32_t d = dst + 16 * ql + hh * (R * 256);
                if (w == 32) {
#pragma unroll
                    for (int r = 0; r < R; ++r) cp_async16_zfill<HINT>(d + r * 256, src + r * 32 * 8, nb, pol_stream);
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r) cp_async16_zfill<HINT>(d + r * 256, src + r * 16 * 8, nb, pol_stream);
                }
            } else if (SPIN2) {
                // 16 rows x 32 columns = 256 16-byte copies: lane (sp, ql) brings the rows of parity sp
                const char *src = half_src + 8 * xcol + (size_t)sp * 8 * w;
                const uint32_t d = dst + 16 * ql + 256 * sp;
                if (w == 64) {
#pragma unroll
                    for (int m = 0; m < R / 2; ++m) cp_async16_zfill<HINT>(d + m * 512, src + m * 2 * 64 * 8, lane_bytes, pol_stream);
                } else if (w == 48) {
#pragma unroll
                    for (int m = 0; m < R / 2; ++m) cp_async16_zfill<HINT>(d + m * 512, src + m * 2 * 48 * 8, lane_bytes, pol_stream);
                } else if (w == 32) {
#pragma unroll
                    for (int m = 0; m < R / 2; ++m) cp_async16_zfill<HINT>(d + m * 512, src + m * 2 * 32 * 8, lane_bytes, pol_stream);
                } else {
#pragma unroll
                    for (int m = 0; m < R / 2; ++m) cp_async16_zfill<HINT>(d + m * 512, src + m * 2 * 16 * 8, lane_bytes, pol_stream);
                }
            } else {
                const char *src = half_src + 16 * lane;
                const uint32_t d = dst + 16 * lane;
                if (w == 64) {
#pragma unroll
                    for (int r = 0; r < HR; ++r) cp_async16_zfill<HINT>(d + r * 512, src + r * 64 * 8, lane_bytes, pol_stream);
                } else if (w == 48) {
#pragma unroll
                    for (int r = 0; r < HR; ++r) cp_async16_zfill<HINT>(d + r * 512, src + r * 48 * 8, lane_bytes, pol_stream);
                } else if (w == 32) {
#pragma unroll
                    for (int r = 0; r < HR; ++r) cp_async16_zfill<HINT>(d + r * 512, src + r * 32 * 8, lane_bytes, pol_stream);
                } else {
#pragma unroll
                    for (int r = 0; r < HR; ++r) cp_async16_zfill<HINT>(d + r * 512, src + r * 16 * 8, lane_bytes, pol_stream);
                }
            }
            stage_w = (stage_w + 1 == STAGES) ? 0 : stage_w + 1;
            --prod_todo;
            if (++prod_half == NH) {   // piece complete: next sub-row (KS: next two)
                prod_half = 0;
                prod_src += (size_t)KSTEP * LAY_ROWS * 8 * w;
                prod_left -= prod_left < KSTEP ? prod_left : KSTEP;
            }
        };
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) {
            if (prod_todo > 0) issue_next();
            if (!full) cp_async_commit();
        }

        int cur_i0 = -1000;
        double2 accA[NS][NI], di[NS][NI];
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int a = 0; a < NI; ++a) accA[s][a] = di[s][a] = make_double2(0.0, 0.0);

        auto flushA = [&]() {
            if (WANTK && cur_i0 > -1000) {
#pragma unroll
                for (int s = 0; s < NS; ++s)
#pragma unroll
                    for (int a = 0; a < NI; ++a) {
                        const int i = cur_i0 + a;
                        const bool ok = lane_in && i >= 0;
                        double *dst = kt_l + s * nl + (size_t)(ok ? i : 0) * ldd + (ok ? l0 : 0);
                        flush_add<FIXED>(ok, dst, accA[s][a].x, fx);
                        flush_add<FIXED>(ok, dst + 1, accA[s][a].y, fx);
                    }
            }
        };

        // reduction helpers (pass h handles values 8h .. 8h+7: value 8h + (lane >> 2), 4 lanes per value, 4 entries each)
        const int accj_rd_off = (lane >> 2) * JK_RED_STRIDE + (lane & 3) * 4;
        // (WIDE: quad lane q reads entries 2q, 2q+1 (+ 8 e) of its value: 128-bit loads, conflict-free with the row stride 40)
        const int red_rd_off = WIDE ? (lane >> 2) * JK_RED_WIDE + (lane & 3) * 2 : accj_rd_off;

        // Operands of a row-block that come from global memory (an L2 round trip takes 1-2 us under the stream's load) are
        // fetched BEFORE the previous row-block's epilogue (drain, flushes):
        //   nx_dj          D_s[j_b, l0..l0+1]                       lane operands, into registers
        //   s_w[r]         D2[i_a, j_b], r = 4 b + a                8-byte cp.async straight into shared memory (registers
        //   s_u[kk][q]     D_s[row_q, k0 + kk], q = s*(4+NB) + (a | 4+b)   held across the epilogue were spilled, and the spill
        //                                                           store waits for the load it was meant to overlap)
        // Row indices are clamped into range: operands of rows outside the canonical range only ever meet stored zeros.
        // (whole-warp chunks: the NV values of sub-row k0 + (lane & 15) are split over the two half-warps, NUH each)
        constexpr int NUH = HALF ? NV : NV / 2;
        static_assert(HALF || NV % 2 == 0, "operand prefetch splits the values over the half-warps");
        double2 nx_dj[NS][NB];
        auto fetch_rb = [&](int rbi) {
            const int imax = rb_imax(rbi), j0 = rb_j0(rbi);
            const int i0 = imax - 3;
            __syncwarp();   // every lane is done with s_w / s_u of the previous sweep
#pragma unroll
            for (int s = 0; s < NS; ++s)
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    nx_dj[s][b] = make_double2(0.0, 0.0);
                    if (WANTK && j0 + b < n && lane_in) nx_dj[s][b] = ldg2(dsp_l + s * nl + (size_t)(j0 + b) * ldd + l0);
                }
            {
                int i = i0 + (lane & 3), j = j0 + ((lane >> 2) & (NB - 1));
                i = i < 0 ? 0 : i;
                j = j < n ? j : n - 1;
                if (lane < R) cp_async8(smem_u32(s_w + lane), p.d2 + (size_t)i * ldd + j);
            }
            if (WANTK) {
                // 32-bit indexing: nspin * n * ldd < 2^31 for every supported n
                const int k = (k0 + (lane & 15)) < n ? (k0 + (lane & 15)) : n - 1;
                const double *dk = dsp_l + k;
                const uint32_t su = smem_u32(s_u + (sp * JK_KN + (lane & 15)) * NVE + (HALF ? 0 : (lane >> 4) * NUH));
                if (!KS || lane < JK_KN) {   // (KS: one density, rows kk = 0..15 staged by the first half-warp)
#pragma unroll
                    for (int qq = 0; qq < NUH; ++qq) {
                        const int q = HALF ? qq : (lane >> 4) * NUH + qq;
                        const int s = q / (NI + NB), wq = q % (NI + NB);
                        int row = wq < NI ? i0 + wq : j0 + (wq - NI);
                        row = row < 0 ? 0 : (row < n ? row : n - 1);
                        cp_async8(su + 8 * qq, dk + (s * n + row) * ldd);
                    }
                }
            }
            cp_async_commit();
        };
        // Density tile rows D2[k0 + kk, l-chunk] (the same 16 rows for every sweep of the task): loaded AHEAD of their sub-row
        // (an L2 hit queues behind the ERI stream and takes longer than one iteration under load) -- two sub-rows for the
        // single-density shapes, one for the dual-density pass (no registers to spare) -- and the last iterations of a
        // sweep fetch the first rows of the NEXT sweep, so no sweep starts by waiting for an L2 round trip.  The loads
        // are unconditional (a predicated load costs a select and a register move that waits for the data at the end of
        // the iteration): rows past the matrix come from the next plane of the work buffer and only meet stored zeros.
        constexpr bool D2_TWO = NB == 4;
        constexpr int D2_AHEAD = D2_TWO ? 2 : 1;
        const double *d2base = p.d2 + (size_t)(k0 + (KS ? hh : 0)) * ldd + (lane_in ? l0 : 0);
        auto load_d2 = [&](int row) { return HINT ? ldg2_keep(d2base + (size_t)row * ldd, pol_keep) : ldg2(d2base + (size_t)row * ldd); };
        // (D2S: the rows go through a two-slot cp.async ring in shared memory instead, one group per row)
        auto copy_d2 = [&](int row, uint32_t slot) {
            cp_async16_zfill<HINT>(smem_u32(s_d2 + slot * 32 + lane), d2base + (size_t)row * ldd, 16u, pol_keep);
            cp_async_commit();
        };
        double2 d2n = make_double2(0.0, 0.0), d2nn = make_double2(0.0, 0.0);
        if (D2S) {
            copy_d2(0, d2slot);
        } else {
            d2n = load_d2(0);
            if (D2_TWO) d2nn = load_d2(KSTEP);
        }
        fetch_rb(0);

        for (int rbi = 0; rbi < task.rb_count; ++rbi) {
            const int imax = rb_imax(rbi), j0 = rb_j0(rbi);
            const int i0 = imax - 3;   // may be negative in the clipped lowest i-block: those rows hold zeros
            const int nkk = imax + 1 - k0 < JK_KN ? imax + 1 - k0 : JK_KN;
            if (i0 != cur_i0) {
                flushA();
                cur_i0 = i0;
#pragma unroll
                for (int s = 0; s < NS; ++s)
#pragma unroll
                    for (int a = 0; a < NI; ++a) {
                        accA[s][a] = make_double2(0.0, 0.0);
                        if (WANTK)
                            di[s][a] = (i0 + a >= 0 && lane_in) ? ldg2(dsp_l + s * nl + (size_t)(i0 + a) * ldd + l0)
                                                                 : make_double2(0.0, 0.0);
                    }
            }
            double2 dj[NS][NB], accB[NS][NB];
#pragma unroll
            for (int s = 0; s < NS; ++s)
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    accB[s][b] = make_double2(0.0, 0.0);
                    dj[s][b] = nx_dj[s][b];
                }
            double accJ[R];
#pragma unroll
            for (int r = 0; r < R; ++r) accJ[r] = 0.0;
            // the prefetched warp-uniform operands s_w / s_u have landed (the group committed by fetch_rb is the newest one:
            // with it every older group -- ring slots of the cp.async path -- is complete as well); the warp barrier at the
            // top of the first iteration makes them visible to all lanes
            cp_async_wait<0>();
            constexpr int NKV = HALF ? NHV : NV;
            constexpr int NKP = KS ? NKV / 4 : (NKV + 1) / 2;

            double cdp[NV];   // lane partials of the previous sub-row, published one iteration later
#pragma unroll
            for (int q = 0; q < NV; ++q) cdp[q] = 0.0;
            const int nit = KS ? (nkk + 1) >> 1 : nkk;         // loop iterations of this sweep

            // one ring slot: wait for the previous half-iteration in every lane (its slot is free, its shared-memory stores are
            // visible), refill the freed slot, wait for the oldest one
            // (D2S: d2row >= 0 = density tile row to fetch for the next iteration; its group is committed BEFORE the ring
            //  slot's, so "all but the newest three groups" = this iteration's row and ring slot are complete)
            auto slot_wait = [&](int d2row) {
                __syncwarp();
                if (D2S && d2row >= 0) copy_d2(d2row, d2slot ^ 1u);
                if (prod_todo > 0) issue_next();
                if (full) {
                    if (D2S) cp_async_wait<1>();
                    mbar_wait(bar_u32 + 8 * stage_r, (phase_bits >> stage_r) & 1u);
                    phase_bits ^= 1u << stage_r;
                } else {
                    cp_async_commit();
                    // this lane's 16 bytes of every row of the oldest slot have landed
                    if (D2S)
                        cp_async_wait<3>();
                    else
                        cp_async_wait<STAGES - 1>();
                }
            };
            auto slot_ready = [&]() -> const double2 * {
                __syncwarp();   // converged from here to the end of the block: no divergent branch below (ptxas can then
                                // schedule the shuffles into the DFMA block instead of fencing them with BRA.DIV)
                const double2 *buf = ring + (size_t)stage_r * (STAGE_BYTES / 16) + (KS ? hh * (R * LW) : 0) + ql;
                stage_r = (stage_r + 1 == STAGES) ? 0 : stage_r + 1;
                return buf;
            };

#pragma unroll 2
            for (int kk = 0; kk < nit; ++kk) {   // (KS: kk counts iterations of two sub-rows, this lane's is kl)
                slot_wait(KSTEP * (kk + 1 < nit ? kk + 1 : 0));   // (row: D2S only; wraps into the next sweep)
                const int kl = KS ? ((2 * kk + hh) < JK_KN ? 2 * kk + hh : JK_KN - 1) : kk;

                // ---- C, loads: partials of sub-row kk-2 (buffer kk & 1; kk-1 and buffer (kk-1) & 1 in the shallow pipeline)
                // (SPIN2: 16 values x 16 entries: value lane >> 1, two lanes per value, 8 entries each)
                // (WIDE: one buffer; its loads stand before the warp barrier of slot_ready(), this iteration's stores after it)
                double r4[NVP][RW], r8[8];
                if (WANTK && WIDE) {
                    const double *rd = s_red + red_rd_off;
#pragma unroll
                    for (int h = 0; h < NVP; ++h)
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const double2 t2 = *reinterpret_cast<const double2 *>(
                                rd + (8 * h < NV - 8 ? 8 * h : (NV > 8 ? NV - 8 : 0)) * JK_RED_WIDE + 8 * e);
                            r4[h][2 * e] = t2.x;
                            r4[h][2 * e + 1] = t2.y;
                        }
                }
                const double2 *buf = slot_ready();
                double rb4[4];   // (half-warp modes with 24 values: values 16 + (lane >> 2), four lanes per value, 4 entries each)
                if (WANTK && HALF) {
                    const double *rb = s_red + ((DEEP ? kk : kk + 1) & 1) * SM::RED_BUF;
                    const double *rd = rb + (lane >> 1) * JK_RED_STRIDE + (lane & 1) * 8;
#pragma unroll
                    for (int e = 0; e < 8; ++e) r8[e] = rd[e];
                    if (NHV > 16) {
                        const double *rd2 = rb + (16 + (lane >> 2)) * JK_RED_STRIDE + (lane & 3) * 4;
#pragma unroll
                        for (int e = 0; e < 4; ++e) rb4[e] = rd2[e];
                    }
                } else if (WANTK && !WIDE) {
                    const double *rd = s_red + ((DEEP ? kk : kk + 1) & 1) * SM::RED_BUF + red_rd_off;
#pragma unroll
                    for (int h = 0; h < NVP; ++h)
#pragma unroll
                        for (int e = 0; e < 4; ++e) r4[h][e] = rd[(8 * h < NV - 8 ? 8 * h : (NV > 8 ? NV - 8 : 0)) * JK_RED_STRIDE + e];
                }
                // ---- A: the FMA block of sub-row kk
                const double2 d2 = D2S ? s_d2[d2slot * 32 + lane] : d2n;
                d2slot ^= 1u;
                if (!D2S) {
                    const int ahead = kk + D2_AHEAD;   // iteration whose row is fetched now (wraps into the next sweep)
                    const double2 nx = load_d2(KSTEP * (ahead < nit ? ahead : ahead - nit));
                    d2n = D2_TWO ? d2nn : nx;
                    if (D2_TWO) d2nn = nx;
                }
                double u[NVE];
                if (WANTK) {
                    const double2 *up = reinterpret_cast<const double2 *>(s_u + (sp * JK_KN + kl) * NVE);
#pragma unroll
                    for (int q = 0; q < NVE / 2; ++q) {
                        const double2 t2 = up[q];
                        u[2 * q] = t2.x;
                        u[2 * q + 1] = t2.y;
                    }
                }
                double2 jt2 = make_double2(0.0, 0.0);
                double cd[NV];
#pragma unroll
                for (int q = 0; q < NV; ++q) cd[q] = 0.0;
                auto fma_rows = [&](const double2 *bf, int half) {
#pragma unroll
                    for (int bb = 0; bb < NB / NH; ++bb)
#pragma unroll
                        for (int a = 0; a < NI; ++a) {
                            const int b = half * (NB / NH) + bb;
                            const int r = 4 * b + a;
                            const double2 x = bf[(4 * bb + a) * LW];
                            accJ[r] = fma(x.x, d2.x, fma(x.y, d2.y, accJ[r]));
                            const double2 w2 = reinterpret_cast<const double2 *>(s_w)[r >> 1];
                            const double wij = (r & 1) ? w2.y : w2.x;
                            jt2.x = fma(x.x, wij, jt2.x);
                            jt2.y = fma(x.y, wij, jt2.y);
                            if (WANTK) {
#pragma unroll
                                for (int s = 0; s < NS; ++s) {
                                    const double ui = u[s * (NI + NB) + a], uj = u[s * (NI + NB) + NI + b];
                                    accA[s][a].x = fma(x.x, uj, accA[s][a].x);   // K~[i,l] += v D[j,k]
                                    accA[s][a].y = fma(x.y, uj, accA[s][a].y);
                                    accB[s][b].x = fma(x.x, ui, accB[s][b].x);   // K~[j,l] += v D[i,k]
                                    accB[s][b].y = fma(x.y, ui, accB[s][b].y);
                                    double &ca = cd[s * (NI + NB) + a], &cb = cd[s * (NI + NB) + NI + b];
                                    ca = fma(x.x, dj[s][b].x, fma(x.y, dj[s][b].y, ca));   // K~[i,k] += v D[j,l]
                                    cb = fma(x.x, di[s][a].x, fma(x.y, di[s][a].y, cb));   // K~[j,k] += v D[i,l]
                                }
                            }
                        }
                };
                fma_rows(buf, 0);
                // ---- C, arithmetic: finish sub-row kk-2 (in the first half's block when the sub-row takes two slots)
                double csum[NVP];
                if (WANTK && HALF) {
                    const int kcol = DEEP ? (kk >= 2 ? kk - 2 : 16) : (kk >= 1 ? kk - 1 : 16);
                    double sum = ((r8[0] + r8[1]) + (r8[2] + r8[3])) + ((r8[4] + r8[5]) + (r8[6] + r8[7]));
                    sum += shfl_xor_f64(sum, 1);
                    s_kout[(lane >> 1) * JK_RED_STRIDE + kcol] = sum;   // both lanes of the pair: same value
                    if (NHV > 16) {
                        double sb = (rb4[0] + rb4[1]) + (rb4[2] + rb4[3]);
                        sb += shfl_xor_f64(sb, 1);
                        sb += shfl_xor_f64(sb, 2);
                        s_kout[(16 + (lane >> 2)) * JK_RED_STRIDE + kcol] = sb;
                    }
                } else if (WANTK) {
#pragma unroll
                    for (int h = 0; h < NVP; ++h) {
                        double sum = (r4[h][0] + r4[h][1]) + (r4[h][2] + r4[h][3]);
                        if (WIDE) sum += (r4[h][RW - 4] + r4[h][RW - 3]) + (r4[h][RW - 2] + r4[h][RW - 1]);
                        sum += shfl_xor_f64(sum, 1);
                        sum += shfl_xor_f64(sum, 2);
                        csum[h] = sum;
                    }
                    // unconditional stores: all four lanes of a quad hold the same sum and store it to the same address;
                    // results that must not land (kk < 2) go to the spare column 16
                    const int kcol = DEEP ? (kk >= 2 ? kk - 2 : 16) : (kk >= 1 ? kk - 1 : 16);
#pragma unroll
                    for (int h = 0; h < NVP; ++h) {
                        const int v = (8 * h < NV - 8 ? 8 * h : (NV > 8 ? NV - 8 : 0)) + (lane >> 2);
                        s_kout[v * JK_RED_STRIDE + kcol] = csum[h];
                    }
                }
                if (NH == 2) {
                    slot_wait(-1);
                    const double2 *buf1 = slot_ready();
                    fma_rows(buf1, 1);
                }
                double2 cur = SLAB ? s_jt[kl * LW + ql] : make_double2(0.0, 0.0);
                // ---- B: publish the partials of sub-row kk-1 (buffer (kk-1) & 1); shallow pipeline: of sub-row kk itself
                if (WANTK) {
                    if (!DEEP) {
#pragma unroll
                        for (int q = 0; q < NV; ++q) cdp[q] = cd[q];
                    }
                    if (!HALF && !WIDE) {   // (half-warp modes: the 16 lanes of a half-warp publish their own partials, nothing
                                            //  to fold; WIDE: all 32 lanes publish)
#pragma unroll
                        for (int q = 0; q < NV; ++q) cdp[q] += shfl_xor_f64(cdp[q], 16);
                    }
                }
                // ---- stores last (they may not be reordered above the loads of this block).  Unconditional: after the xor
                // shuffle both half-warps hold the same value and store it to the same address.
                cur.x += jt2.x;
                cur.y += jt2.y;
                if (SLAB) {
                    s_jt[kl * LW + ql] = cur;     // (SPIN2: both half-warps hold the same J~ contributions; KS: own rows)
                } else {
                    // unconditional: cells with l > k hold exact zeros (stored zeros times finite operands); lanes past the
                    // matrix add their zeros to column 0
                    double *dst = p.jt + (size_t)(k0 + kl) * ldd + (lane_in ? l0 : 0);
                    flush_add<FIXED>(true, dst, cur.x, fx);
                    flush_add<FIXED>(true, dst + 1, cur.y, fx);
                }
                if (WANTK) {
                    double *wr = WIDE ? s_red + lane
                                      : s_red + ((DEEP ? kk + 1 : kk) & 1) * SM::RED_BUF + hh * (NV * JK_RED_STRIDE) + (lane & 15);
#pragma unroll
                    for (int q = 0; q < NV; ++q) wr[q * SM::RED_STRIDE] = cdp[q];
                    if (DEEP) {
#pragma unroll
                        for (int q = 0; q < NV; ++q) cdp[q] = cd[q];
                    }
                }
            }
            if (D2_TWO && !D2S && nit == 1) d2n = load_d2(0);   // (one-iteration sweep: its look-ahead of two skipped the next row 0)
            if (rbi + 1 < task.rb_count) fetch_rb(rbi + 1);   // in flight during the epilogue below
            {   // ---- epilogue of the row-block
            // (RBS: the row-block's scalars are re-read from shared memory here -- kept live across the sweep they were spilled)
            const int imax_pre = imax, j0_pre = j0;
            const int imax = RBS ? reinterpret_cast<const volatile RowBlock *>(s_rb)[rbi].imax : imax_pre;
            const int j0 = RBS ? reinterpret_cast<const volatile RowBlock *>(s_rb)[rbi].j0 : j0_pre;
            const int i0 = imax - 3;
            const int nkk = imax + 1 - k0 < JK_KN ? imax + 1 - k0 : JK_KN;
            const int nit = KS ? (nkk + 1) >> 1 : nkk;
            // ---- drain the reduction pipeline: sub-rows nkk-2 (published, not summed) and nkk-1 (not yet published)
            if (WANTK) {
                __syncwarp();
                if (DEEP) {
                    if (HALF) {
                        double *wr = s_red + ((nit + 1) & 1) * SM::RED_BUF + hh * (NV * JK_RED_STRIDE) + ql;
#pragma unroll
                        for (int q = 0; q < NV; ++q) wr[q * JK_RED_STRIDE] = cdp[q];
                    } else {
#pragma unroll
                        for (int q = 0; q < NV; ++q) cdp[q] += shfl_xor_f64(cdp[q], 16);
                        if (lane < 16) {
                            double *wr = s_red + ((nit + 1) & 1) * SM::RED_BUF + lane;
#pragma unroll
                            for (int q = 0; q < NV; ++q) wr[q * JK_RED_STRIDE] = cdp[q];
                        }
                    }
                    __syncwarp();
                }
#pragma unroll
                for (int d = DEEP ? 0 : 1; d < 2; ++d) {   // d = 0: iteration nit-2 (buffer nit & 1), d = 1: iteration nit-1
                    const int ks = nit - 2 + d;
                    if (HALF) {
                        const double *rb = s_red + ((nit + d) & 1) * SM::RED_BUF;
                        const double *rd = rb + (lane >> 1) * JK_RED_STRIDE + (lane & 1) * 8;
                        double sum = ((rd[0] + rd[1]) + (rd[2] + rd[3])) + ((rd[4] + rd[5]) + (rd[6] + rd[7]));
                        sum += shfl_xor_f64(sum, 1);
                        if ((lane & 1) == 0 && ks >= 0) s_kout[(lane >> 1) * JK_RED_STRIDE + ks] = sum;
                        if (NHV > 16) {
                            const double *rd2 = rb + (16 + (lane >> 2)) * JK_RED_STRIDE + (lane & 3) * 4;
                            double sb = (rd2[0] + rd2[1]) + (rd2[2] + rd2[3]);
                            sb += shfl_xor_f64(sb, 1);
                            sb += shfl_xor_f64(sb, 2);
                            if ((lane & 3) == 0 && ks >= 0) s_kout[(16 + (lane >> 2)) * JK_RED_STRIDE + ks] = sb;
                        }
                        continue;
                    }
                    const double *rd = s_red + (WIDE ? 0 : ((nit + d) & 1) * SM::RED_BUF) + red_rd_off;
#pragma unroll
                    for (int h = 0; h < NVP; ++h) {
                        const int v = 8 * h + (lane >> 2);
                        double sum = 0.0;
                        if (WIDE) {
                            if (v < NV) {
                                const double *rv = rd + 8 * h * JK_RED_WIDE;
                                sum = ((rv[0] + rv[1]) + (rv[8] + rv[9])) + ((rv[16] + rv[17]) + (rv[24] + rv[25]));
                            }
                        } else if (v < NV)
                            sum = (rd[8 * h * JK_RED_STRIDE] + rd[8 * h * JK_RED_STRIDE + 1]) +
                                  (rd[8 * h * JK_RED_STRIDE + 2] + rd[8 * h * JK_RED_STRIDE + 3]);
                        sum += shfl_xor_f64(sum, 1);
                        sum += shfl_xor_f64(sum, 2);
                        if ((lane & 3) == 0 && v < NV && ks >= 0) s_kout[v * JK_RED_STRIDE + ks] = sum;
                    }
                }
                __syncwarp();
                // reduced K~[row_q, k0 .. k0+nkk) leave with coalesced reds (16 consecutive k per value): lane = kk + 16 (q & 1),
                // pass m covers q = 2m, 2m+1
                // (SPIN2: 16 values, value V = 8 spin + q, pass m covers V = 2m, 2m+1.  KS: 16 values V = 8 parity + q in 8
                //  iteration columns: lane = it + 8 (V & 3), pass m covers V = 4m .. 4m+3, sub-row kk = 2 it + (V >> 3))
                // (the destinations are computed here, not before the sweep: live across the loop they were spilled, and a
                //  local-memory reload is an L2 round trip with the whole L1 carved out as shared memory)
#pragma unroll
                for (int m = 0; m < NKP; ++m) {
                    const int qv = KS ? 4 * m + (lane >> 3) : 2 * m + (lane >> 4);
                    const int q = HALF ? qv % NV : qv, par = HALF ? qv / NV : 0;   // par: spin (SPIN2) / sub-row parity (KS)
                    const int s = SPIN2 ? par : q / (NI + NB), wq = q % (NI + NB);
                    const int row = wq < NI ? i0 + wq : j0 + (wq - NI);
                    const int kkq = KS ? 2 * (lane & 7) + par : (lane & 15);
                    const bool ok = qv < NKV && row >= 0 && row < n && kkq < nkk;
                    const double v = qv < NKV ? s_kout[qv * JK_RED_STRIDE + (KS ? (lane & 7) : (lane & 15))] : 0.0;
                    flush_add<FIXED>(ok, p.kt + (ok ? s * nl + (size_t)row * ldd + k0 + kkq : 0), v, fx);
                }
#pragma unroll
                for (int s = 0; s < NS; ++s)
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        const bool ok = lane_in && (j0 + b < n);
                        double *dst = kt_l + s * nl + (size_t)(ok ? j0 + b : 0) * ldd + (ok ? l0 : 0);
                        flush_add<FIXED>(ok, dst, accB[s][b].x, fx);
                        flush_add<FIXED>(ok, dst + 1, accB[s][b].y, fx);
                    }
            }
            {
                // J~[i_a, j_b]: reduce the R row dots across the warp (same shared-memory transpose, 8 per pass)
                if (!SPIN2) {   // (SPIN2: both half-warps hold the same row dots over the chunk's 16 column pairs)
#pragma unroll
                    for (int r = 0; r < R; ++r) accJ[r] += shfl_xor_f64(accJ[r], 16);
                }
#pragma unroll
                for (int h = 0; h < (R + 7) / 8; ++h) {
                    __syncwarp();
                    if (lane < 16) {
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            if (8 * h + q < R) s_red[q * JK_RED_STRIDE + lane] = accJ[8 * h + q];
                    }
                    __syncwarp();
                    const int v = 8 * h + (lane >> 2);
                    const double *rd = s_red + accj_rd_off;
                    double sum = 0.0;
                    if (v < R) sum = (rd[0] + rd[1]) + (rd[2] + rd[3]);
                    sum += shfl_xor_f64(sum, 1);
                    sum += shfl_xor_f64(sum, 2);
                    const int i = i0 + (v & 3), j = j0 + (v >> 2);
                    const bool ok = (lane & 3) == 0 && v < R && i >= 0 && j <= i;   // other slots hold zeros
                    flush_add<FIXED>(ok, p.jt + (ok ? (size_t)i * ldd + j : 0), sum, fx);
                }
            }
            }
        }
        flushA();
        __syncwarp();
        int imax0 = RBS ? s_rb[lane].imax : my_imax;   // longest sweep of the task (lanes without a row-block hold k0)
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const int other = __shfl_xor_sync(0xffffffffu, imax0, o);
            imax0 = other > imax0 ? other : imax0;
        }
        if (SLAB && lane_in && sp == 0) {
            const int kmax = imax0 + 1 - k0 < JK_KN ? imax0 + 1 - k0 : JK_KN;
            for (int kk = KS ? hh : 0; kk < kmax; kk += KSTEP) {   // (KS: the half-warp that owns the row's parity)
                const int k = k0 + kk;
                const double2 cur = s_jt[kk * LW + ql];
                flush_add<FIXED>(l0 <= k, p.jt + (size_t)k * ldd + l0, cur.x, fx);
                flush_add<FIXED>(l0 + 1 <= k, p.jt + (size_t)k * ldd + l0 + 1, cur.y, fx);
            }
        }
        __syncwarp();
    }
}

}  // namespace sqcc
