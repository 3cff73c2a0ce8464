// Fused J/K kernel over the native v2 layout (layout.cuh).  Included by jk_fused.cu.
//
// A WARP is the unit of work.  A warp task is (16 sub-rows k) x (one 64-column chunk, two columns per lane) x (a list
// of row-blocks: 4 rows i x NB rows j of one tile, NB = 4 for the single-density passes, 2 for the dual-density UHF
// pass).  In the v2 layout the R = 4 NB rows of a row-block for one (k, chunk) are ONE contiguous piece and the 16
// sub-rows of a task follow each other, so the ERI stream of a task is a single running pointer.  A slot of the warp's
// shared-memory ring holds the piece of one loop iteration followed by the density tile row D2[k, chunk] of that sub-row:
//   * full pieces (64 columns): one elected lane issues two cp.async.bulk (TMA) per slot -- piece and density row -- on
//     one per-slot mbarrier (UBLKCP in SASS);
//   * the triangular edge (pieces of 48 columns in this kernel, of 16 / 32 columns in the narrow-piece instantiation that
//     runs two sub-rows per iteration on the half-warps): 16-byte cp.async per lane with immediate row offsets and the
//     zero-fill form for the lanes past the piece, landing in the same slot layout.
// Cells outside the canonical range are stored as 0.0, so the consumer is mask free and has no per-row address
// arithmetic at all.
//
// Consumer, per sub-row: every lane holds D[j,l], D[i,l] (its two columns), the accumulators K~[i,l], K~[j,l], J~[i,j];
// D[i,k], D[j,k] and D2[i,j] are warp-uniform (staged per row-block in shared memory by cp.async issued before the previous
// row-block's epilogue, broadcast loads); J~[k,l] goes to a warp-private slab that lives for the whole task.  K~[i,k],
// K~[j,k] need a sum over the lanes once per sub-row; that reduction is SOFTWARE-PIPELINED over three iterations so that
// the loop body is one straight-line block in which ptxas can place all the shuffles / shared-memory round trips in the
// issue slots the half-rate DFMAs leave free:
//     iteration kk:   C  loads of the partials of kk-2 (before the warp barrier), 3 adds, 2 shuffles, one store -> s_kout[.][kk-2]
//                     A  the 12 R DFMAs of sub-row kk                                                     -> lane partials cd(kk)
//                     B  the half-warps swap halves of cd(kk-1) (one xor-16 shuffle per value pair) and store -> s_red
// and the 16 x NV reduced values of a row-block leave with coalesced red.global.add.f64 after its sweep.  (The
// dual-density pass has no registers for the third stage: it publishes cd(kk) itself, all 32 lane partials, unshuffled.)
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
// xor-shuffle of a double without the volatile mov.b64 pack/unpack of the CUDA header
__device__ __forceinline__ double shfl_xor_f64(double v, int lane_mask)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xffffffffu, lo, lane_mask);
    hi = __shfl_xor_sync(0xffffffffu, hi, lane_mask);
    return __hiloint2double(hi, lo);
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
constexpr int JK_D2_BYTES = 512;   // density tile row(s) that ride with every ring slot: one double2 per lane

// Shared memory of one warp.
// SPIN2 (HM = 1, dual-density UHF pass): the two half-warps work on the SAME integrals (16 rows x 32 columns,
// two columns per lane) for the two spin densities, so every per-spin array exists once per lane (NS = 1) and the per-spin
// shared-memory arrays twice per warp.
// KS (HM = 2, narrow pieces of the triangular edge, widths 16 / 32): the two half-warps work on two CONSECUTIVE SUB-ROWS
// of the same 32 columns, so a loop iteration -- whose cost is its 12 R DFMAs whatever the piece width -- covers two
// sub-rows instead of one.
// A ring slot = the piece(s) of one loop iteration [R rows][64 | 2 x 32 columns] followed by the density tile row(s)
// D2[k, chunk] of the same sub-row(s) (one double2 per lane): the row an iteration needs arrives with its integrals (same
// mbarrier / cp.async group) instead of through a register load that has to be issued sub-rows ahead and rotated.
template <int NB, int NS, int STAGES, int HM>
struct JKSmem {
    static constexpr bool SPIN2 = HM == 1, KS = HM == 2, HALF = HM != 0;
    static constexpr int R = 4 * NB;
    static constexpr int NV = NS * (4 + NB);          // lane-partial values reduced per sub-row (per half-warp if HALF)
    static constexpr int NVE = (NV + 1) & ~1;
    static constexpr int NSP = SPIN2 ? 2 : 1;                // spin planes staged in s_u
    static constexpr int NRV = HALF ? 2 * NV : NV;           // values reduced per iteration (both half-warps)
    static constexpr int ROW_BYTES = HALF ? 256 : 512;       // one row of a ring slot / of the J~ slab
    static constexpr int STAGE_BYTES = (KS ? 2 : 1) * R * ROW_BYTES;     // integrals of one slot
    static constexpr int STAGE_STRIDE = STAGE_BYTES + JK_D2_BYTES;
    // WIDE_RED (4 x 2 dual-density pass, whole-warp chunks): the reduction scratch keeps all 32 lane partials of a value (no
    // shuffle before the store): rows of 32 doubles, odd rows stored with their column xor 8, which makes the 128-bit reads
    // of two neighbouring rows by one quarter-warp conflict free without padding
    static constexpr bool WIDE_RED = NB == 2 && HM == 0;
    // doubles (>= 8 values x 17: J~ row dots go 8 at a time)
    static constexpr int RED_BUF = WIDE_RED ? NRV * 32 : (NRV > 8 ? NRV : 8) * JK_RED_STRIDE;
    static constexpr int OFF_SLAB = 0;                                   // double2 [16][32 | 16]
    static constexpr int OFF_RING = OFF_SLAB + JK_KN * ROW_BYTES;        // [STAGES][STAGE_STRIDE]
    static constexpr int OFF_U = OFF_RING + STAGES * STAGE_STRIDE;       // double [NSP][16][NVE]
    static constexpr int OFF_RED = OFF_U + NSP * JK_KN * NVE * 8;        // double [RED_BUF] (ONE buffer: the loads of an
                                                                         // iteration stand before the warp barrier its stores follow)
    static constexpr int OFF_KOUT = OFF_RED + RED_BUF * 8;               // double [NRV][17]
    static constexpr int OFF_W = OFF_KOUT + ((NRV + 1) & ~1) * JK_RED_STRIDE * 8;     // double [R]
    static constexpr int OFF_BAR = OFF_W + R * 8;                        // u64 [STAGES]
    // RBS (dual-density pass): the row-block list of the task lives in shared memory instead of one entry per lane +
    // shuffles -- the pass has no registers to spare, and a spilled register is an L2 round trip
    static constexpr bool RBS = NB == 2 && HM == 0;
    static constexpr int OFF_RB = (OFF_BAR + STAGES * 8 + 15) & ~15;     // RowBlock [32]
    static constexpr int WARP_BYTES = (OFF_RB + (RBS ? 512 : 0) + 63) & ~63;
};

// NB: rows j per row-block (4: RHF / J-only, 2: UHF).  TMA: full pieces through cp.async.bulk (false: cp.async only).
// HINT: L2 cache-policy hints (evict-first on the ERI stream, evict-last on the density tile rows).
// HM = 1 (SPIN2): dual-density pass with the spins on the two half-warps (NB = 4, cp.async ring, 32-column chunks: task.chunk
// counts 32-column chunks); lane = 16 spin + column pair.
// HM = 2 (KS): narrow-piece pass (tasks whose pieces are 16 or 32 columns wide), half-warp = sub-row parity.
template <int NB, int NSPIN, bool WANTK, int WARPS, int STAGES, bool TMA, bool FIXED, bool HINT, int HM>
__global__ void __launch_bounds__(WARPS * 32, 1) jk_stream_kernel(const JKParams p)
{
    constexpr int NI = 4;
    constexpr bool SPIN2 = HM == 1, KS = HM == 2, HALF = HM != 0;
    constexpr int NS = (WANTK && !SPIN2) ? NSPIN : 1;
    using SM = JKSmem<NB, NS, STAGES, HM>;
    constexpr int R = SM::R, NV = SM::NV, NVE = SM::NVE;
    static_assert(!SPIN2 || (NB == 4 && !TMA && WANTK && NSPIN == 2), "spin-split pass: 4 x 4 row-blocks, cp.async ring");
    static_assert(!KS || !TMA, "narrow-piece pass: cp.async ring");
    constexpr int LW = HALF ? 16 : 32;                // double2 per ring-slot row = lanes that share one row
    constexpr int KSTEP = KS ? 2 : 1;                 // sub-rows per loop iteration
    constexpr int NVP = (NV + 7) / 8;
    constexpr int NHV = 2 * NV;   // half-warp modes: values reduced per iteration, 16 lane partials each (16, or 24 dual-density)
    static_assert(!HALF || NHV == 16 || NHV == 24, "half-warp reduction: one pass of 16 values, a second one of 8");
    constexpr bool DEEP = NB == 4;   // reduction pipeline over three iterations (needs NV more live registers) or two
    constexpr int NVH = NV / 2;      // whole-warp chunks: values a half-warp publishes (see "publish" below)
    constexpr bool WIDE = SM::WIDE_RED;   // 32 lane partials per value in the scratch (JKSmem)
    constexpr int RW = WIDE ? 8 : 4;      // scratch entries one lane sums per value
    static_assert(!WIDE || (!DEEP && !HALF && NV > 8 && (NV - 8) % 2 == 0), "wide reduction rows: shallow pipeline, whole-warp chunks");
    constexpr uint32_t STAGE_BYTES = SM::STAGE_BYTES, STAGE_STRIDE = SM::STAGE_STRIDE;
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
    constexpr bool RBS = SM::RBS;
    const uint32_t ring_u32 = smem_u32(wbase + SM::OFF_RING);
    const uint32_t bar_u32 = smem_u32(wbase + SM::OFF_BAR);
    const int n = p.n, ldd = p.ldd;
    const size_t nl = (size_t)n * ldd;
    const double fx = p.fx_scale;
    const double *dsp_l = p.dsp + (size_t)sp * nl;    // density / accumulator plane of this lane's spin
    double *kt_l = p.kt + (size_t)sp * nl;
    const uint64_t pol_stream = HINT ? l2_policy_evict_first() : 0, pol_keep = HINT ? l2_policy_evict_last() : 0;
    uint32_t stage_w = 0, stage_r = 0, phase_bits = 0;   // ring write / read positions and mbarrier parities, persist across tasks

    // the narrow-piece pass lets the kernel launched behind it (same accumulators, no dependence on this one's results)
    // take over every SM the moment this grid's CTA leaves it
    if (KS) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
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

#pragma unroll 4
        for (int kk = 0; kk < JK_KN; ++kk) s_jt[kk * LW + ql] = make_double2(0.0, 0.0);

        // row-block list of the task, one entry per lane (rb_count <= 32).  The tasks that share a group of row-blocks
        // (other k-ranges / chunks) are neighbours in the queue and run at the same time on other warps; every task starts
        // at its own offset into the (cyclic) list so that they do not all flush the same rows j at the same moment.
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
        // loop iterations (ring slots) of this task
        int total_it = 0;
        {
            int cnt = lane < task.rb_count ? (my_imax + 1 - k0 < JK_KN ? my_imax + 1 - k0 : JK_KN) : 0;
            if (KS) cnt = (cnt + 1) >> 1;   // two sub-rows per iteration
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            total_it = cnt;
        }

        // ---- producer: one ring slot (the piece(s) of one iteration + their density tile row(s)) per call
        int prod_rbi = -1, prod_left = 0, prod_todo = total_it;
        const char *prod_src = reinterpret_cast<const char *>(p.P);
        // D2[k0 + (KS: parity), chunk]: the rows are the same for every row-block of the task.  Lanes past the matrix read
        // column 0 (their integrals are stored zeros); rows past the sweep / the matrix come from the next plane of the work
        // buffer and only meet stored zeros or are never used.
        const char *const d2_row0 = reinterpret_cast<const char *>(p.d2 + (size_t)k0 * ldd + cbase);
        const char *prod_d2 = d2_row0;
        const int d2_lane_off = 8 * ((KS ? hh * ldd : 0) + (lane_in ? xcol : -cbase));
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
                prod_d2 = d2_row0;
            }
            const uint32_t dst = ring_u32 + stage_w * STAGE_STRIDE;
            if (full) {
                if (lane == 0) {
                    const uint32_t bar = bar_u32 + 8 * stage_w;
                    mbar_expect_tx(bar, STAGE_BYTES + JK_D2_BYTES);
                    if (HINT) {
                        bulk_g2s_hint(dst, prod_src, STAGE_BYTES, bar, pol_stream);
                        bulk_g2s_hint(dst + STAGE_BYTES, prod_d2, JK_D2_BYTES, bar, pol_keep);
                    } else {
                        bulk_g2s(dst, prod_src, STAGE_BYTES, bar);
                        bulk_g2s(dst + STAGE_BYTES, prod_d2, JK_D2_BYTES, bar);
                    }
                }
            } else {
                if (KS) {
                    // two consecutive narrow pieces, one per half-warp: lane (hh, ql) brings its own 16 bytes of the R rows of
                    // sub-row 2 it + hh (zero-filled past the piece width and past the last sub-row of the sweep)
                    const uint32_t nb = (xcol < w && hh < prod_left) ? 16u : 0u;
                    const char *src = prod_src + 16 * ql + (size_t)(hh < prod_left ? hh : 0) * (LAY_ROWS * 8) * w;
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
                    const char *src = prod_src + 8 * xcol + (size_t)sp * 8 * w;
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
                    const char *src = prod_src + 16 * lane;
                    const uint32_t d = dst + 16 * lane;
                    if (w == 64) {
#pragma unroll
                        for (int r = 0; r < R; ++r) cp_async16_zfill<HINT>(d + r * 512, src + r * 64 * 8, lane_bytes, pol_stream);
                    } else if (w == 48) {
#pragma unroll
                        for (int r = 0; r < R; ++r) cp_async16_zfill<HINT>(d + r * 512, src + r * 48 * 8, lane_bytes, pol_stream);
                    } else if (w == 32) {
#pragma unroll
                        for (int r = 0; r < R; ++r) cp_async16_zfill<HINT>(d + r * 512, src + r * 32 * 8, lane_bytes, pol_stream);
                    } else {
#pragma unroll
                        for (int r = 0; r < R; ++r) cp_async16_zfill<HINT>(d + r * 512, src + r * 16 * 8, lane_bytes, pol_stream);
                    }
                }
                // this lane's 16 bytes of the density tile row (it reads back only what it copied itself)
                cp_async16_zfill<HINT>(dst + STAGE_BYTES + 16 * lane, prod_d2 + d2_lane_off, 16u, pol_keep);
            }
            stage_w = (stage_w + 1 == STAGES) ? 0 : stage_w + 1;
            --prod_todo;
            prod_src += (size_t)KSTEP * LAY_ROWS * 8 * w;   // next sub-row (KS: next two)
            prod_d2 += (size_t)KSTEP * 8 * ldd;
            prod_left -= prod_left < KSTEP ? prod_left : KSTEP;
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
        const int red_rd_off = (lane >> 2) * JK_RED_STRIDE + (lane & 3) * 4;
        // Publishing the lane partials of a sub-row (whole-warp chunks): the two half-warps SWAP halves of their values with
        // one xor-16 shuffle per value pair -- the lower half-warp ends up with values 0 .. NVH-1 summed over lanes (L, L+16),
        // the upper one with NVH .. NV-1 -- and every lane stores NVH values: 16 entries per value in the scratch, each
        // written once (half the shuffles and half the stores of folding all NV values onto one half-warp; the shared-memory
        // data pipe is ~70 % busy in this kernel, so wavefronts count).
        // (WIDE: no shuffle, all 32 lanes store their NV partials; quad lane q later reads entries 2q, 2q+1 (+ 8 e) of its value)
        const bool upper = lane >= 16;
        const int wide_rd_a = (lane >> 2) * 32 + (lane & 3) * 2 + 8 * ((lane >> 2) & 1), wide_rd_b = wide_rd_a ^ 8;
        auto publish = [&](const double *cdv) {
            if (WIDE) {
#pragma unroll
                for (int q = 0; q < NV; ++q) s_red[q * 32 + (lane ^ ((q & 1) * 8))] = cdv[q];
            } else {
                double keep[NVH > 0 ? NVH : 1];
#pragma unroll
                for (int q = 0; q < NVH; ++q) {
                    const double mine = upper ? cdv[q + NVH] : cdv[q], give = upper ? cdv[q] : cdv[q + NVH];
                    keep[q] = mine + shfl_xor_f64(give, 16);
                }
                double *wr = s_red + (upper ? NVH : 0) * JK_RED_STRIDE + (lane & 15);
#pragma unroll
                for (int q = 0; q < NVH; ++q) wr[q * JK_RED_STRIDE] = keep[q];
            }
        };

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

            // one ring slot: wait for the previous iteration in every lane (its slot is free, its shared-memory stores are
            // visible), refill the freed slot, wait for the oldest one
            auto slot_wait = [&]() {
                __syncwarp();
                if (prod_todo > 0) issue_next();
                if (full) {
                    mbar_wait(bar_u32 + 8 * stage_r, (phase_bits >> stage_r) & 1u);
                    phase_bits ^= 1u << stage_r;
                } else {
                    cp_async_commit();
                    cp_async_wait<STAGES - 1>();   // this lane's 16 bytes of every row of the oldest slot have landed
                }
            };
            const double2 *d2src = ring;
            auto slot_ready = [&]() -> const double2 * {
                __syncwarp();   // converged from here to the end of the block: no divergent branch below (ptxas can then
                                // schedule the shuffles into the DFMA block instead of fencing them with BRA.DIV)
                const double2 *slot = ring + (size_t)stage_r * (STAGE_STRIDE / 16);
                d2src = slot + STAGE_BYTES / 16 + lane;
                stage_r = (stage_r + 1 == STAGES) ? 0 : stage_r + 1;
                return slot + (KS ? hh * (R * LW) : 0) + ql;
            };

#pragma unroll 2
            for (int kk = 0; kk < nit; ++kk) {   // (KS: kk counts iterations of two sub-rows, this lane's is kl)
                slot_wait();
                const int kl = KS ? ((2 * kk + hh) < JK_KN ? 2 * kk + hh : JK_KN - 1) : kk;

                // ---- C, loads: lane partials the previous iteration published (of sub-row kk-2; kk-1 in the shallow
                // pipeline).  They stand BEFORE the warp barrier of slot_ready(), this iteration's stores to the same (single)
                // buffer after it.
                // (half-warp modes: NHV values x 16 entries: value lane >> 1, two lanes per value, 8 entries each; with 24
                //  values a second pass: value 16 + (lane >> 2), four lanes per value, 4 entries each)
                double r4[NVP][RW], r8[8], rb4[4];
                if (WANTK && HALF) {
                    const double *rd = s_red + (lane >> 1) * JK_RED_STRIDE + (lane & 1) * 8;
#pragma unroll
                    for (int e = 0; e < 8; ++e) r8[e] = rd[e];
                    if (NHV > 16) {
                        const double *rd2 = s_red + (16 + (lane >> 2)) * JK_RED_STRIDE + (lane & 3) * 4;
#pragma unroll
                        for (int e = 0; e < 4; ++e) rb4[e] = rd2[e];
                    }
                } else if (WANTK && WIDE) {
#pragma unroll
                    for (int h = 0; h < NVP; ++h)
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            // (the second pass starts at the even row NV - 8: same column swizzle)
                            const double2 t2 = *reinterpret_cast<const double2 *>(
                                s_red + (h == 0 ? 0 : NV - 8) * 32 + ((e & 1) ? wide_rd_b : wide_rd_a) + 16 * (e >> 1));
                            r4[h][2 * e] = t2.x;
                            r4[h][2 * e + 1] = t2.y;
                        }
                } else if (WANTK) {
                    const double *rd = s_red + red_rd_off;
#pragma unroll
                    for (int h = 0; h < NVP; ++h)
#pragma unroll
                        for (int e = 0; e < 4; ++e) r4[h][e] = rd[(8 * h < NV - 8 ? 8 * h : (NV > 8 ? NV - 8 : 0)) * JK_RED_STRIDE + e];
                }
                const double2 *buf = slot_ready();
                // ---- A: the FMA block of sub-row kk
                const double2 d2 = *d2src;   // D2[k0 + kl, l0..l0+1]: arrived with the slot
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
#pragma unroll
                for (int b = 0; b < NB; ++b)
#pragma unroll
                    for (int a = 0; a < NI; ++a) {
                        const int r = 4 * b + a;
                        const double2 x = buf[r * LW];
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
                // ---- C, arithmetic: finish the sub-row whose partials were loaded above.  Unconditional stores: all lanes
                // of a group hold the same sum and store it to the same address; results that must not land (first
                // iterations) go to the spare column 16
                const int kcol = DEEP ? (kk >= 2 ? kk - 2 : 16) : (kk >= 1 ? kk - 1 : 16);
                if (WANTK && HALF) {
                    double sum = ((r8[0] + r8[1]) + (r8[2] + r8[3])) + ((r8[4] + r8[5]) + (r8[6] + r8[7]));
                    sum += shfl_xor_f64(sum, 1);
                    s_kout[(lane >> 1) * JK_RED_STRIDE + kcol] = sum;
                    if (NHV > 16) {
                        double sb = (rb4[0] + rb4[1]) + (rb4[2] + rb4[3]);
                        sb += shfl_xor_f64(sb, 1);
                        sb += shfl_xor_f64(sb, 2);
                        s_kout[(16 + (lane >> 2)) * JK_RED_STRIDE + kcol] = sb;
                    }
                } else if (WANTK) {
                    double csum[NVP];
#pragma unroll
                    for (int h = 0; h < NVP; ++h) {
                        double sum = (r4[h][0] + r4[h][1]) + (r4[h][2] + r4[h][3]);
                        if (WIDE) sum += (r4[h][RW - 4] + r4[h][RW - 3]) + (r4[h][RW - 2] + r4[h][RW - 1]);
                        sum += shfl_xor_f64(sum, 1);
                        sum += shfl_xor_f64(sum, 2);
                        csum[h] = sum;
                    }
#pragma unroll
                    for (int h = 0; h < NVP; ++h) {
                        const int v = (8 * h < NV - 8 ? 8 * h : (NV > 8 ? NV - 8 : 0)) + (lane >> 2);
                        s_kout[v * JK_RED_STRIDE + kcol] = csum[h];
                    }
                }
                double2 cur = s_jt[kl * LW + ql];
                // ---- B: publish the partials of sub-row kk-1; shallow pipeline: of sub-row kk itself.  Stores last (they may
                // not be reordered above the loads of this block).
                cur.x += jt2.x;
                cur.y += jt2.y;
                s_jt[kl * LW + ql] = cur;     // (SPIN2: both half-warps hold the same J~ contributions; KS: own rows)
                if (WANTK) {
                    if (!DEEP) {
#pragma unroll
                        for (int q = 0; q < NV; ++q) cdp[q] = cd[q];
                    }
                    if (HALF) {   // the 16 lanes of a half-warp publish their own partials (values hh NV .. hh NV + NV - 1)
                        double *wr = s_red + hh * (NV * JK_RED_STRIDE) + (lane & 15);
#pragma unroll
                        for (int q = 0; q < NV; ++q) wr[q * JK_RED_STRIDE] = cdp[q];
                    } else {
                        publish(cdp);
                    }
                    if (DEEP) {
#pragma unroll
                        for (int q = 0; q < NV; ++q) cdp[q] = cd[q];
                    }
                }
            }
            if (rbi + 1 < task.rb_count) fetch_rb(rbi + 1);   // in flight during the epilogue below
            {   // ---- epilogue of the row-block
            // (RBS: the row-block's scalars are re-read from shared memory here -- kept live across the sweep they were spilled)
            const int imax_pre = imax, j0_pre = j0;
            const int imax = RBS ? reinterpret_cast<const volatile RowBlock *>(s_rb)[rbi].imax : imax_pre;
            const int j0 = RBS ? reinterpret_cast<const volatile RowBlock *>(s_rb)[rbi].j0 : j0_pre;
            const int i0 = imax - 3;
            const int nkk = imax + 1 - k0 < JK_KN ? imax + 1 - k0 : JK_KN;
            const int nit = KS ? (nkk + 1) >> 1 : nkk;
            // ---- drain the reduction pipeline: the scratch holds the partials the last iteration published (of iteration
            // nit-2, deep pipeline, or nit-1); the deep pipeline still has those of iteration nit-1 in registers
            if (WANTK) {
                auto drain_reduce = [&](int ks) {   // sum the published partials into s_kout[.][ks]
                    if (HALF) {
                        const double *rd = s_red + (lane >> 1) * JK_RED_STRIDE + (lane & 1) * 8;
                        double sum = ((rd[0] + rd[1]) + (rd[2] + rd[3])) + ((rd[4] + rd[5]) + (rd[6] + rd[7]));
                        sum += shfl_xor_f64(sum, 1);
                        if ((lane & 1) == 0 && ks >= 0) s_kout[(lane >> 1) * JK_RED_STRIDE + ks] = sum;
                        if (NHV > 16) {
                            const double *rd2 = s_red + (16 + (lane >> 2)) * JK_RED_STRIDE + (lane & 3) * 4;
                            double sb = (rd2[0] + rd2[1]) + (rd2[2] + rd2[3]);
                            sb += shfl_xor_f64(sb, 1);
                            sb += shfl_xor_f64(sb, 2);
                            if ((lane & 3) == 0 && ks >= 0) s_kout[(16 + (lane >> 2)) * JK_RED_STRIDE + ks] = sb;
                        }
                        return;
                    }
                    const double *rd = s_red + red_rd_off;
#pragma unroll
                    for (int h = 0; h < NVP; ++h) {
                        const int v = 8 * h + (lane >> 2);
                        double sum = 0.0;
                        if (WIDE) {
                            if (v < NV) {
                                const double *ra = s_red + 8 * h * 32 + wide_rd_a, *rb = s_red + 8 * h * 32 + wide_rd_b;
                                sum = ((ra[0] + ra[1]) + (rb[0] + rb[1])) + ((ra[16] + ra[17]) + (rb[16] + rb[17]));
                            }
                        } else if (v < NV)
                            sum = (rd[8 * h * JK_RED_STRIDE] + rd[8 * h * JK_RED_STRIDE + 1]) +
                                  (rd[8 * h * JK_RED_STRIDE + 2] + rd[8 * h * JK_RED_STRIDE + 3]);
                        sum += shfl_xor_f64(sum, 1);
                        sum += shfl_xor_f64(sum, 2);
                        if ((lane & 3) == 0 && v < NV && ks >= 0) s_kout[v * JK_RED_STRIDE + ks] = sum;
                    }
                };
                __syncwarp();
                drain_reduce(DEEP ? nit - 2 : nit - 1);
                if (DEEP) {
                    __syncwarp();
                    if (HALF) {
                        double *wr = s_red + hh * (NV * JK_RED_STRIDE) + ql;
#pragma unroll
                        for (int q = 0; q < NV; ++q) wr[q * JK_RED_STRIDE] = cdp[q];
                    } else {
                        publish(cdp);
                    }
                    __syncwarp();
                    drain_reduce(nit - 1);
                }
                __syncwarp();
                // reduced K~[row_q, k0 .. k0+nkk) leave with coalesced reds (16 consecutive k per value): lane = kk + 16 (q & 1),
                // pass m covers q = 2m, 2m+1
                // (SPIN2: 16 values, value V = 8 spin + q, pass m covers V = 2m, 2m+1.  KS: NHV values V = NV parity + q in 8
                //  iteration columns: lane = it + 8 (V & 3), pass m covers V = 4m .. 4m+3, sub-row kk = 2 it + parity)
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
                    const double *rd = s_red + red_rd_off;
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
        if (lane_in && sp == 0) {
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
    // launched with programmatic stream serialization behind the narrow-piece pass, this grid may run out of tasks while a
    // CTA of that pass is still working: it must not complete (and release the kernel that reads the accumulators)
    // before the pass has.  A no-op for a plain launch.
    if (!KS) asm volatile("griddepcontrol.wait;" ::: "memory");
}

}  // namespace sqcc
