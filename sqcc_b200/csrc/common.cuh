// Shared declarations for libsqcc_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/sqcc_b200.h"
#include "layout.cuh"

namespace sqcc {

// ---------------------------------------------------------------- error plumbing
void set_error(const std::string &msg);
#define SQCC_CUDA(call)                                                                              \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            sqcc::set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                     \
            return SQCC_ERR_CUDA;                                                                    \
        }                                                                                            \
    } while (0)
#define SQCC_REQUIRE(cond, msg)                                                                      \
    do {                                                                                             \
        if (!(cond)) {                                                                               \
            sqcc::set_error(std::string(msg));                                                       \
            return SQCC_ERR_INVALID;                                                                 \
        }                                                                                            \
    } while (0)
#define SQCC_TRY(expr)                                                                               \
    do {                                                                                             \
        int s_ = (expr);                                                                             \
        if (s_ != SQCC_OK) return s_;                                                                \
    } while (0)

// ---------------------------------------------------------------- packed layout arithmetic (native v2: layout.cuh)
__host__ __device__ __forceinline__ int64_t pair_index(int64_t i, int64_t j) { return i * (i + 1) / 2 + j; }

// ---------------------------------------------------------------- J/K tile schedule
constexpr int JK_KN = 16;      // k-range per warp task
constexpr int JK_CHUNK = 64;   // l-chunk per warp task (2 per lane, one 128-bit load) == LAY_CHUNK

struct RowBlock {        // 4 rows i x NB rows j of one tile
    int64_t base;        // offset (doubles, shard-relative) of the tile
    int imax;            // last i of the tile's i-block (rows i = imax-3 .. imax)
    int j0;              // first j (4 jb, or 4 jb + 2 for the second 4 x 2 half)
};
struct JKTask {
    int k0;        // first k of the range (multiple of JK_KN)
    int chunk;     // l-chunk index (l0 = 64 * chunk)
    int rb_begin;  // first row-block (index into the row-block list)
    int rb_count;
};
struct JKSchedule {
    int nb = 0;          // rows j per row-block (4 or 2)
    std::vector<RowBlock> h_rowblocks;
    std::vector<JKTask> h_tasks;
    RowBlock *d_rowblocks = nullptr;
    JKTask *d_tasks = nullptr;
    int n_tasks = 0;
};

// Peer group of the fused J/K exchange (comm.cu)
constexpr int COMM_MAXW = 16;
struct CommRank;
struct Comm {
    int world = 1, rank = 0;
    bool attached = false;      // peer_base[0..world) are valid
    bool enabled = true;        // sqcc_comm_set_enabled: false -> sqcc_jk_build returns partials even with a group attached
    bool ipc = false;           // peers were opened with cudaIpcOpenMemHandle (other processes)
    bool same_device = false;   // in-process group whose contexts share one device (single-launch emulation)
    double *d_sym = nullptr;    // symmetric region [acc 3 nl][res 3 nl][flags 2 COMM_MAXW u64]
    void *peer_base[COMM_MAXW] = {};
    unsigned long long epoch = 0;
    unsigned int bar_count = 0;
    unsigned int *d_gridbar = nullptr;
    unsigned long long *d_stamps = nullptr;   // [8] phase timestamps of the last exchange launch
    CommRank *d_view = nullptr;  // [COMM_MAXW]
    bool view_valid = false;     // d_view[0] holds this rank's view for outputs (view_J, view_K)
    double *view_J = nullptr, *view_K = nullptr;
    int *h_status = nullptr, *d_status = nullptr;   // mapped host word
};

struct ScfState;   // scf.cu

// CUDA-event timer with a small ring of event pairs: back-to-back calls (a benchmark's timed loop) each get their own
// pair, sqcc_get_timings returns the MEAN over the pairs recorded since the previous query (at most TIMER_RING).
constexpr int TIMER_RING = 64;
struct Timer {
    cudaEvent_t a[TIMER_RING] = {}, b[TIMER_RING] = {};
    int next = 0;     // pair the next start() records into
    int count = 0;    // pairs recorded since the last query (capped at TIMER_RING)
    bool used = false;
};

}  // namespace sqcc

struct sqcc_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    int64_t launches = 0;

    // ERI shard (caller-owned): the contiguous tile range [t_begin, t_end) of the native v2 tensor (layout.cuh).
    double *d_eri = nullptr;
    int n = 0;
    int ldd = 0;  // padded leading dimension of the D/J~/K~ work matrices (multiple of 4)
    int64_t t_begin = 0, t_end = 0;
    int64_t shard_base = 0;      // offset (doubles) of tile t_begin in the full tensor
    std::vector<int64_t> h_ibase, h_tfirst;   // [nb + 1] offset / tile index of the first tile of every i-block
    int64_t *d_ibase = nullptr, *d_tfirst = nullptr;   // device copies (one allocation)
    std::vector<int> jlo, jhi;   // [n] rows held: for every i the interval of j (host-side description of the shard)
    int64_t nrows = 0;           // held rows
    bool whole = false;          // every tile of the tensor is held (unsharded)
    int i_begin = 0, i_end = 0;  // i range with held rows
    int64_t native_len = 0;
    int64_t n_unique = 0;
    double eri_absmax = -1.0;    // max |stored value| (lazily computed: deterministic fixed-point mode)
    sqcc::JKSchedule sched[7];    // [0]: 4x4 row-blocks, all tasks  [1]: 4x2 (UHF)  [2]: 4x4, 32-column chunks (spin-split UHF)
                                  // [3]: 4x4, tasks with 48 / 64-column pieces  [4]: 4x4, narrow tasks (16 / 32 columns)
                                  // [5], [6]: the same split of the 4x2 tasks
    int jk_variant = 0;
    int jk_cfg = 0;  // tuning knob of the streaming kernel, see jk_partial
    int jk_fixed = 0;  // deterministic mode: 64-bit fixed-point accumulation (sqcc_jk_set_deterministic)
    double fx_scale = 0.0;   // fixed-point scale of the accumulators of the last build (0: floating point)

    // J/K scratch
    double *d_dwork = nullptr;   // [3][N][ldd]: D2 = 2*(Da+Db), Da, Db (padded)
    double *d_acc = nullptr;     // [3][N][ldd]: J~, K~a, K~b (first part of comm.d_sym)
    sqcc::Comm comm;
    unsigned int *d_counter = nullptr;
    double *d_io = nullptr;      // device staging for *_host calls
    size_t io_bytes = 0;
    double *h_pin = nullptr;     // pinned staging
    size_t pin_bytes = 0;

    // grid
    const double *d_phi = nullptr;
    const double *d_w = nullptr;
    int64_t G = 0;
    int grid_n = 0;
    double *d_gridwork = nullptr;
    size_t gridwork_bytes = 0;

    // mp2 scratch
    double *d_mp2work = nullptr;
    size_t mp2work_bytes = 0;
    double *d_mp2out = nullptr;
    size_t mp2out_n = 0;

    sqcc::ScfState *scf = nullptr;   // device-resident SCF iteration state (scf.cu)

    sqcc::Timer timers[8];
};

namespace sqcc {
int ensure_io(sqcc_ctx *ctx, size_t bytes);
int ensure_pin(sqcc_ctx *ctx, size_t bytes);
int timer_start(sqcc_ctx *ctx, int slot);
int timer_stop(sqcc_ctx *ctx, int slot);

// eri_layout.cu
void pair_unrank_host(int64_t ij, int &i, int &j);
int64_t tile_count(int n);
void tile_tables(int n, std::vector<int64_t> &ibase, std::vector<int64_t> &tfirst);
int64_t tile_offset(int n, int64_t t);
int tile_intervals(int n, int64_t t_begin, int64_t t_end, std::vector<int> &jlo, std::vector<int> &jhi);
int eri_build_tables(sqcc_ctx *ctx);
int eri_synth(sqcc_ctx *ctx, uint64_t seed);
int eri_pack_full_slab(sqcc_ctx *ctx, const double *d_slab, int p0, int np);
int eri_pack_canonical(sqcc_ctx *ctx, const double *d_rows, int64_t ij0, int64_t nrows);
int eri_unpack_full(sqcc_ctx *ctx, double *d_full);
int eri_pack_blocks(sqcc_ctx *ctx, const double *d_values, const int64_t *d_desc, int nblk);

// jk_fused.cu
int jk_build_schedule(sqcc_ctx *ctx);
int jk_partial(sqcc_ctx *ctx, const double *d_D, int nspin, int want_k);
int jk_build(sqcc_ctx *ctx, const double *d_D, int nspin, int want_k, double *d_J, double *d_K);

// comm.cu
int comm_alloc(sqcc_ctx *ctx);
void comm_free(sqcc_ctx *ctx);
int jk_reduce(sqcc_ctx *ctx, int nspin, int want_k, double *d_J, double *d_K);

// grid_lda.cu
int lda_vxc(sqcc_ctx *ctx, const double *d_D, int nspin, double *d_Vxc, double *d_Exc, double *d_rho);
int grid_quadrature(sqcc_ctx *ctx, const double *d_v, double *d_V);
int grid_rho(sqcc_ctx *ctx, const double *d_D, double *d_rho);

// scf.cu
void scf_free(sqcc_ctx *ctx);

// mp2.cu
int ao2mo(sqcc_ctx *ctx, const double *d_C1, int n1, const double *d_C2, int n2, const double *d_C3, int n3,
          const double *d_C4, int n4, double *d_out);
int mp2(sqcc_ctx *ctx, const double *d_C, const double *d_eps, int n_occ, double *d_ecorr, double *d_iajb);
}  // namespace sqcc
