// pgcn_b200.cu — C-ABI implementation (see include/pgcn_b200.h for the contract and the
// reference file:line each entry point replaces).
//
// Host side of the hot path: plan construction (device CSR copies, row-block schedule, boundary
// CSR for the gradient scatter-add), kernel dispatch, and the two transports of the halo
// exchange (NCCL grouped send/recv resolved at run time from the process's libnccl, and the
// peer-memory path that stores rows directly into the neighbour GPU's slab over NVLink).
#include "../../include/pgcn_b200.h"
#include "spmm_kernels.cuh"
#include "spmm_ring.cuh"

#include <dlfcn.h>
#include <unistd.h>
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace pgcn;

// ------------------------------------------------------------------------------------------
// NCCL, resolved lazily from whatever libnccl the process already has (torch's bundled copy),
// so this library has no link-time NCCL dependency and loads on a box without it.
// ------------------------------------------------------------------------------------------
namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess_ = 0, ncclFloat_ = 7 };

struct NcclApi {
    bool tried = false, ok = false;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

bool load_nccl()
{
    if (g_nccl.tried) return g_nccl.ok;
    g_nccl.tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // torch's copy, if loaded
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return false;
#define PGCN_SYM(field, name)                                              \
    *(void**)(&g_nccl.field) = dlsym(h, name);                             \
    if (!g_nccl.field) return false;
    PGCN_SYM(GetUniqueId, "ncclGetUniqueId")
    PGCN_SYM(CommInitRank, "ncclCommInitRank")
    PGCN_SYM(CommDestroy, "ncclCommDestroy")
    PGCN_SYM(Send, "ncclSend")
    PGCN_SYM(Recv, "ncclRecv")
    PGCN_SYM(GroupStart, "ncclGroupStart")
    PGCN_SYM(GroupEnd, "ncclGroupEnd")
    PGCN_SYM(GetErrorString, "ncclGetErrorString")
#undef PGCN_SYM
    g_nccl.ok = true;
    return true;
}

std::string g_lib_error = "";

constexpr int kMaxPeers = 16;

struct DevCsr {
    int nrows = 0;                  // rows of the matrix (output rows)
    int nrows_c = 0;                // non-empty rows = rows the schedule walks
    int64_t nnz = 0;
    std::vector<int> h_rowptr;      // COMPACT row pointers (empty rows squeezed out), for scheduling
    int* d_cw = nullptr;            // entries in 272-byte pieces of 32: col[32] | val[32] | row-end mask | cold mask | pad
    int* d_rowids = nullptr;        // compact row -> output row, null when the identity
    int* d_empty = nullptr;         // output rows without entries (zero-filled when beta == 0)
    int nempty = 0;
    bool view = false;              // a row range of another matrix: d_cw / d_rowids / d_empty are borrowed
    int row_base = 0;               // compact id of the first walked row (views: offset into the parent's numbering)
    std::vector<int> h_rowids, h_empty;   // host copies (parents of views only)
    unsigned char* d_final = nullptr;     // per walked row: this matrix is the LAST launch of the forward that writes it
    int64_t tuned_epb[2] = {0, 0};  // per-matrix autotune results (0: use the plan option): [0] register, [1] ring
    int tuned_slots = 0;
    // schedules: [0] register-pipeline kernel (small row blocks, one per lane group),
    //            [1] shared-memory ring kernel (large row blocks, one per warp)
    struct Sched {
        int4* d_blocks = nullptr;
        int nblocks = 0;
        int4* d_long = nullptr;
        int nlong = 0;
        int nslots = 0;
        float* d_partial = nullptr;
        int64_t epb = -1, long_row = -1;
    } sched[2];
};

struct P2PBlob {                     // what pgcn_p2p_export writes (PGCN_P2P_HANDLE_BYTES)
    cudaIpcMemHandle_t ipc;          // 64 bytes
    int64_t arena_bytes;
    int64_t off_flags, off_fwd[2], off_bwd[2];
    int32_t k, rank, f_max, pad;
    int64_t pid;                     // exporting process: a peer in the same process is reached through local_ptr
    void* local_ptr;
    int64_t send_off[kMaxPeers + 1];
    int64_t recv_off[kMaxPeers + 1];
};
static_assert(sizeof(P2PBlob) <= PGCN_P2P_HANDLE_BYTES, "blob too large");

}  // namespace

struct pgcn_plan {
    int device = 0;
    int m = 0, h = 0, k = 1, rank = 0, f_max = 0;
    int64_t S = 0;
    DevCsr fwd, tr;                  // A_local over [own | halo] columns, and its transpose
    // split A_local = [A_own | A_halo(peer 0) | ... ] (Parallel-GCN/main.c:271 then :295 per received block):
    DevCsr own;                      // columns < m
    std::vector<DevCsr> halo_q;      // per source peer: boundary rows x that peer's columns (slab-relative)
    DevCsr tr_own;                   // view: rows [0, m) of tr
    std::vector<DevCsr> tr_halo_q;   // views: rows of tr that belong to each peer
    bool have_split = false;
    int64_t cols_ref = 0, rows_ref_t = 0;

    std::vector<int64_t> send_off, recv_off;
    int* d_send_idx = nullptr;
    // boundary CSR for unpack_add
    int* d_brow = nullptr; int* d_bptr = nullptr; int* d_bpos = nullptr; int nb = 0;

    // slabs (local transport)
    float* d_send_slab = nullptr;    // S x f_max
    float* d_halo_slab = nullptr;    // h x f_max   (forward receive)
    float* d_rrecv_slab = nullptr;   // S x f_max   (reverse receive)
    float* d_hsend_slab = nullptr;   // h x f_max   (reverse send: halo partials of A^T g)

    // options
    int64_t opt_epb = 128, opt_long = 0, opt_tile = 0, opt_overlap = 1, opt_hot_mb = 64;
    int64_t opt_relu = 0;            // fused layer epilogue of pgcn_forward: Z = max(0, A_local * H)
    // kernel: 0 auto (ring with TMA bulk copies where it applies), 4 register pipeline, 5 ring/1-D TMA,
    //         6 ring/cp.async, 7 ring/TMA tile::gather4 (= auto)
    int64_t opt_ring_groups = 2;
    int64_t opt_persistent_multi = 0;
    int64_t opt_kernel = 0, opt_ring_slots = 16, opt_ring_epb = 512, opt_ring_long = 0, opt_persistent = 1;
    bool ring_attr_set[48] = {false};
    int ring_ctas_per_sm[48] = {0};
    unsigned int* d_counter = nullptr;     // block counters of the persistent ring kernel (one per feature tile)
    int num_sms = 148;

    // NCCL
    ncclComm_t comm = nullptr;
    bool comm_borrowed = false;            // pgcn_comm_share: the communicator belongs to another plan
    cudaStream_t comm_stream = nullptr;
    cudaStream_t host_stream = nullptr;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    std::vector<cudaEvent_t> ev_step;      // one per peer step of the pipelined exchange
    unsigned int* d_done = nullptr;        // per-destination CTA counters of put_rows_kernel

    // peer-memory transport
    bool p2p = false;
    int64_t opt_p2p = 1;                   // 0: never use the peer transport (all ranks must agree)
    void* arena = nullptr; int64_t arena_bytes = 0;
    int64_t off_flags = 0, off_fwd[2] = {0, 0}, off_bwd[2] = {0, 0};
    void* peer_arena[kMaxPeers] = {nullptr};
    bool peer_local[kMaxPeers] = {false};   // peer lives in THIS process (same-process plans: no IPC mapping)
    P2PBlob peer_blob[kMaxPeers];
    unsigned long long epoch = 0;

    // host-buffer variant: two device slots, copy-in / compute / copy-out streams chained by events
    float* d_hostH[2] = {nullptr, nullptr}; float* d_hostZ[2] = {nullptr, nullptr}; int64_t host_cap = 0;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    int64_t host_steps = 0;

    int64_t launches = 0;
    std::string err;
};

namespace {

int fail(pgcn_plan* p, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (p) p->err = buf; else g_lib_error = buf;
    return code;
}

#define CU(p, expr)                                                                         \
    do {                                                                                    \
        cudaError_t e__ = (expr);                                                           \
        if (e__ != cudaSuccess)                                                             \
            return fail(p, PGCN_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                   \
                        cudaGetErrorString(e__), __FILE__, __LINE__);                       \
    } while (0)

#define NC(p, expr)                                                                         \
    do {                                                                                    \
        int e__ = (expr);                                                                   \
        if (e__ != ncclSuccess_)                                                            \
            return fail(p, PGCN_ERR_NCCL, "%s failed: %s", #expr, g_nccl.GetErrorString(e__)); \
    } while (0)

template <class T>
int upload(pgcn_plan* p, T** dst, const T* src, size_t n)
{
    *dst = nullptr;
    CU(p, cudaMalloc((void**)dst, std::max<size_t>(n, 1) * sizeof(T)));
    if (n) CU(p, cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

// Upload one CSR. Rows without entries are squeezed out of the walked row space (their outputs are
// zero-filled by a separate launch); `ext_rowmap` maps the rows of an already-compact matrix (the
// halo-column part, which only holds boundary rows) to output rows.
// `col_refs[j]` = number of stored entries in column j and `cold_thresh` the reference count at or
// below which a column is COLD (-1: no marking): cold columns get kColdFlag and are gathered with an
// L2 evict_first policy, the most-referenced rows of H (as many as fit the hot budget) evict_last.
int csr_upload(pgcn_plan* p, DevCsr& c, int nrows, const int* rowptr, const int* colidx, const float* vals,
               const std::vector<int>* ext_rowmap = nullptr, const int* col_refs = nullptr, int cold_thresh = -1,
               bool keep_host = false)
{
    c.nrows = nrows;
    c.nnz = rowptr[nrows];
    const size_t npieces = (size_t)((c.nnz + 31) / 32 + 1);
    std::vector<int> cw(npieces * kPieceInts, 0);
    for (int64_t e = 0; e < c.nnz; ++e) {
        int* pc = cw.data() + (size_t)(e >> 5) * kPieceInts;
        const int i = (int)(e & 31);
        pc[i] = colidx[e];
        memcpy(&pc[32 + i], &vals[e], 4);
        if (col_refs && cold_thresh >= 0 && col_refs[colidx[e]] <= cold_thresh) pc[65] |= (int)(1u << i);
    }
    std::vector<int> rowids, empty;
    c.h_rowptr.clear();
    c.h_rowptr.push_back(0);
    for (int r = 0; r < nrows; ++r) {
        if (rowptr[r + 1] > rowptr[r]) {
            { const int64_t e = (int64_t)rowptr[r + 1] - 1; cw[(size_t)(e >> 5) * kPieceInts + 64] |= (int)(1u << (e & 31)); }
            c.h_rowptr.push_back(rowptr[r + 1]);
            rowids.push_back(ext_rowmap ? (*ext_rowmap)[r] : r);
        } else if (!ext_rowmap) {
            empty.push_back(r);
        }
    }
    c.nrows_c = (int)rowids.size();
    c.nempty = (int)empty.size();
    int rc;
    if ((rc = upload(p, &c.d_cw, cw.data(), cw.size()))) return rc;
    if (ext_rowmap || c.nempty > 0) {
        if ((rc = upload(p, &c.d_rowids, rowids.data(), rowids.size()))) return rc;
    }
    if (c.nempty > 0) {
        if ((rc = upload(p, &c.d_empty, empty.data(), empty.size()))) return rc;
    }
    if (keep_host) { c.h_rowids.swap(rowids); c.h_empty.swap(empty); }
    return 0;
}

// Rows [r0, r1) of an uploaded matrix as a matrix of its own (no copy): the transposed CSR's own rows and the
// rows that belong to one peer are contiguous row ranges, so the pipelined backward needs no second copy of A^T.
void csr_view(const DevCsr& base, DevCsr& v, int r0, int r1)
{
    v = DevCsr();
    v.view = true;
    int c0 = r0, c1 = r1, e0 = 0, e1 = 0;
    if (base.d_rowids) {                     // rows were squeezed: compact ids of the range
        c0 = (int)(std::lower_bound(base.h_rowids.begin(), base.h_rowids.end(), r0) - base.h_rowids.begin());
        c1 = (int)(std::lower_bound(base.h_rowids.begin(), base.h_rowids.end(), r1) - base.h_rowids.begin());
        e0 = (int)(std::lower_bound(base.h_empty.begin(), base.h_empty.end(), r0) - base.h_empty.begin());
        e1 = (int)(std::lower_bound(base.h_empty.begin(), base.h_empty.end(), r1) - base.h_empty.begin());
    }
    v.nrows = r1 - r0;
    v.nrows_c = c1 - c0;
    v.row_base = c0;
    v.h_rowptr.assign(base.h_rowptr.begin() + c0, base.h_rowptr.begin() + c1 + 1);   // absolute entry offsets
    v.nnz = (int64_t)v.h_rowptr.back() - v.h_rowptr.front();
    v.d_cw = base.d_cw;
    v.d_rowids = base.d_rowids;
    v.d_empty = base.d_empty ? base.d_empty + e0 : nullptr;
    v.nempty = e1 - e0;
}

void csr_free(DevCsr& c)
{
    if (!c.view) { cudaFree(c.d_cw); cudaFree(c.d_rowids); cudaFree(c.d_empty); }
    cudaFree(c.d_final);
    for (auto& sc : c.sched) { cudaFree(sc.d_blocks); cudaFree(sc.d_long); cudaFree(sc.d_partial); }
    c = DevCsr();
}

// Cut the (compact) row range into row blocks of about `epb` nnz (at most kMaxRowsPerBlock rows); rows longer
// than `long_row` become ceil(deg/epb) single-row segments with a slot each in the side buffer.
constexpr int kMaxRowsPerBlock = 128;

// Pure host function (also reachable through pgcn_debug_schedule for CPU-side tests).
void make_schedule(const int* rp, int nrows_c, int64_t epb, int64_t long_row,
                   std::vector<int4>& blocks, std::vector<int4>& longs, int& nslots,
                   int max_rows = kMaxRowsPerBlock, int row_base = 0)
{
    blocks.clear(); longs.clear(); nslots = 0;
    const int64_t nnz = nrows_c > 0 ? rp[nrows_c] : 0;
    blocks.reserve((size_t)(nnz / epb + nrows_c / kMaxRowsPerBlock + 16));
    int cur_begin = 0;         // first row of the open block
    int64_t cur_edges = 0;
    auto close = [&](int row_end) {
        if (row_end > cur_begin)
            blocks.push_back(make_int4(row_base + cur_begin, row_end - cur_begin, rp[cur_begin], rp[row_end]));
        cur_begin = row_end;
        cur_edges = 0;
    };
    for (int r = 0; r < nrows_c; ++r) {
        const int64_t d = (int64_t)rp[r + 1] - rp[r];
        if (d > long_row) {
            close(r);
            const int nseg = (int)((d + epb - 1) / epb);
            longs.push_back(make_int4(row_base + r, nslots, nseg, 0));
            for (int s = 0; s < nseg; ++s) {
                const int e0 = rp[r] + (int)(s * epb);
                const int e1 = (int)std::min<int64_t>((int64_t)rp[r + 1], (int64_t)e0 + epb);
                blocks.push_back(make_int4(row_base + r, -(nslots + 1), e0, e1));
                ++nslots;
            }
            cur_begin = r + 1;
            continue;
        }
        if (cur_edges > 0 && cur_edges + d > epb) close(r);
        cur_edges += d;
        if (r + 1 - cur_begin >= max_rows) close(r + 1);
    }
    close(nrows_c);
}

int build_schedule(pgcn_plan* p, DevCsr& c, int which, int64_t epb, int64_t long_row)
{
    DevCsr::Sched& sc = c.sched[which];
    if (sc.epb == epb && sc.long_row == long_row) return 0;

    std::vector<int4> blocks, longs;
    int nslots = 0;
    make_schedule(c.h_rowptr.data(), c.nrows_c, epb, long_row, blocks, longs, nslots,
                  which == 1 ? (1 << 30) : kMaxRowsPerBlock, c.row_base);

    cudaFree(sc.d_blocks); cudaFree(sc.d_long); cudaFree(sc.d_partial);
    sc.d_blocks = nullptr; sc.d_long = nullptr; sc.d_partial = nullptr;
    int rc;
    if ((rc = upload(p, &sc.d_blocks, blocks.data(), blocks.size()))) return rc;
    if ((rc = upload(p, &sc.d_long, longs.data(), longs.size()))) return rc;
    CU(p, cudaMalloc((void**)&sc.d_partial, std::max<size_t>((size_t)nslots * p->f_max, 1) * sizeof(float)));
    sc.nblocks = (int)blocks.size();
    sc.nlong = (int)longs.size();
    sc.nslots = nslots;
    sc.epb = epb;
    sc.long_row = long_row;
    return 0;
}

// ---- kernel dispatch ---------------------------------------------------------------------

struct TileCfg { int lpe, vpl, vw, tiles; };

int pow2ceil(int x) { int p = 1; while (p < x) p <<= 1; return p; }

TileCfg choose_tile(const pgcn_plan* p, int f)
{
    TileCfg t;
    t.vw = (f % 4 == 0) ? 4 : 1;
    const int nvec = f / t.vw;
    int tile_vecs = nvec;
    if (p->opt_tile > 0) tile_vecs = std::max<int>(1, (int)std::min<int64_t>(nvec, p->opt_tile / t.vw));
    tile_vecs = std::min(tile_vecs, 128);
    t.lpe = std::max(4, std::min(32, pow2ceil(tile_vecs)));
    int vpl = (tile_vecs + t.lpe - 1) / t.lpe;
    t.vpl = vpl <= 1 ? 1 : (vpl <= 2 ? 2 : 4);
    t.tiles = (nvec + t.lpe * t.vpl - 1) / (t.lpe * t.vpl);
    return t;
}

typedef void (*spmm_fn)(const SpmmArgs);

template <int LPE, int VW>
spmm_fn pick_vpl(int vpl, bool halo)
{
    switch (vpl) {
        case 1: return halo ? spmm_rowblock_kernel<LPE, 1, VW, true> : spmm_rowblock_kernel<LPE, 1, VW, false>;
        case 2: return halo ? spmm_rowblock_kernel<LPE, 2, VW, true> : spmm_rowblock_kernel<LPE, 2, VW, false>;
        default: return halo ? spmm_rowblock_kernel<LPE, 4, VW, true> : spmm_rowblock_kernel<LPE, 4, VW, false>;
    }
}

template <int VW>
spmm_fn pick_lpe(int lpe, int vpl, bool halo)
{
    switch (lpe) {
        case 4: return pick_vpl<4, VW>(vpl, halo);
        case 8: return pick_vpl<8, VW>(vpl, halo);
        case 16: return pick_vpl<16, VW>(vpl, halo);
        default: return pick_vpl<32, VW>(vpl, halo);
    }
}

typedef void (*ring_fn)(const SpmmArgs, const RingArgs);
typedef void (*ring_g4_fn)(const SpmmArgs, const RingArgs, const CUtensorMap, const CUtensorMap);

// cuTensorMapEncodeTiled, resolved through the runtime (no link-time dependency on libcuda)
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
encode_tiled_fn g_encode_tiled = nullptr;
bool g_encode_tried = false;

encode_tiled_fn encode_tiled()
{
    if (!g_encode_tried) {
        g_encode_tried = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            g_encode_tiled = reinterpret_cast<encode_tiled_fn>(fn);
        else
            cudaGetLastError();
    }
    return g_encode_tiled;
}

// Tensor map of a row-major fp32 matrix [rows, f] for tile::gather4 loads of `tile` floats per row.
bool make_row_map(CUtensorMap* tm, const float* base, int64_t rows, int f, int tile, int box_rows)
{
    encode_tiled_fn enc = encode_tiled();
    if (!enc || !base) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)f, (cuuint64_t)std::max<int64_t>(rows, 1)};
    const cuuint64_t gstride[1] = {(cuuint64_t)f * 4};
    const cuuint32_t box[2] = {(cuuint32_t)tile, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Ring kernel instances: shape id = 0: G=8,NG=2 (16 slots)  1: G=16,NG=2 (32)  2: G=32,NG=2 (64)  3: G=16,NG=4 (64)
struct RingShape { int g, ng; };
const RingShape kRingShapes[4] = {{8, 2}, {16, 2}, {32, 2}, {16, 4}};

template <int VPL, bool HALO>
ring_fn pick_ring_t(int shape, int mode)
{
    if (mode == 1) return spmm_ring_kernel<VPL, 8, 2, 1, HALO>;
    return shape == 1 ? spmm_ring_kernel<VPL, 16, 2, 0, HALO> : spmm_ring_kernel<VPL, 8, 2, 0, HALO>;
}
ring_fn pick_ring(int vpl, int shape, int mode, bool halo)
{
    if (vpl == 2) return halo ? pick_ring_t<2, true>(shape, mode) : pick_ring_t<2, false>(shape, mode);
    return halo ? pick_ring_t<1, true>(shape, mode) : pick_ring_t<1, false>(shape, mode);
}
template <int VPL, bool HALO>
ring_g4_fn pick_ring_g4_t(int shape)
{
    switch (shape) {
        case 1: return spmm_ring_g4_kernel<VPL, 16, 2, HALO>;
        case 2: return spmm_ring_g4_kernel<VPL, 32, 2, HALO>;
        case 3: return spmm_ring_g4_kernel<VPL, 16, 4, HALO>;
        default: return spmm_ring_g4_kernel<VPL, 8, 2, HALO>;
    }
}
ring_g4_fn pick_ring_g4(int vpl, int shape, bool halo)
{
    if (vpl == 2) return halo ? pick_ring_g4_t<2, true>(shape) : pick_ring_g4_t<2, false>(shape);
    return halo ? pick_ring_g4_t<1, true>(shape) : pick_ring_g4_t<1, false>(shape);
}

// CUDA loads kernels lazily, at their first launch, and that load synchronises with the device. A rank whose
// stream already holds a spinning p2p_wait_kernel must therefore never launch a not-yet-loaded kernel behind it
// when the ranks it waits for live in the SAME process (single-process multi-rank use: tests, smoke) — their put
// kernels would never be enqueued. Touching every kernel once, when the peer transport is set up, removes the hazard.
template <class F>
void touch_kernel(F fn) { cudaFuncAttributes fa; if (cudaFuncGetAttributes(&fa, (const void*)fn) != cudaSuccess) cudaGetLastError(); }

void preload_kernels()
{
    static bool done = false;
    if (done) return;
    done = true;
    for (int halo = 0; halo < 2; ++halo)
        for (int lpe = 4; lpe <= 32; lpe *= 2)
            for (int vpl = 1; vpl <= 4; vpl *= 2) { touch_kernel(pick_lpe<4>(lpe, vpl, halo != 0)); touch_kernel(pick_lpe<1>(lpe, vpl, halo != 0)); }
    for (int vpl = 1; vpl <= 2; ++vpl)
        for (int halo = 0; halo < 2; ++halo) {
            for (int shape = 0; shape < 4; ++shape) touch_kernel(pick_ring_g4(vpl, shape, halo != 0));
            for (int shape = 0; shape < 2; ++shape) touch_kernel(pick_ring(vpl, shape, 0, halo != 0));
            touch_kernel(pick_ring(vpl, 0, 1, halo != 0));
        }
    touch_kernel(zero_rows_kernel<4>); touch_kernel(zero_rows_kernel<1>);
    touch_kernel(spmm_fixup_kernel<4>); touch_kernel(spmm_fixup_kernel<1>);
    touch_kernel(pack_rows_kernel<4>); touch_kernel(pack_rows_kernel<1>);
    touch_kernel(unpack_add_kernel<4>); touch_kernel(unpack_add_kernel<1>);
    touch_kernel(put_rows_kernel<4>); touch_kernel(p2p_wait_kernel);
}

bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

// Which SpMM kernel serves width f with these operands: the shared-memory ring (TMA bulk copies) needs whole
// 128-float vectors and 16-byte aligned rows; everything else takes the register pipeline.
bool use_ring(const pgcn_plan* p, const float* H0, const float* H1, int f)
{
    if (p->opt_kernel == 4) return false;
    return f % 128 == 0 && aligned16(H0) && aligned16(H1);
}

int launch_spmm(pgcn_plan* p, DevCsr& c, const float* H0, const float* H1, int split,
                float* Z0, float* Z1, int zsplit, int f, int beta, cudaStream_t st, int relu = 0, bool use_final = false)
{
    if (c.nrows == 0) return 0;
    const bool ring = use_ring(p, H0, H1, f) && aligned16(Z0) && aligned16(Z1);
    int64_t epb, long_row;
    if (ring) {
        epb = std::max<int64_t>(c.tuned_epb[1] > 0 ? c.tuned_epb[1] : p->opt_ring_epb, 64);
        long_row = p->opt_ring_long > 0 ? p->opt_ring_long : 2 * epb;
    } else {
        epb = std::max<int64_t>(c.tuned_epb[0] > 0 ? c.tuned_epb[0] : p->opt_epb, 8);
        long_row = p->opt_long > 0 ? p->opt_long : 4 * epb;
    }
    int rc = build_schedule(p, c, ring ? 1 : 0, epb, long_row);
    if (rc) return rc;
    const DevCsr::Sched& sc = c.sched[ring ? 1 : 0];
    const TileCfg t = choose_tile(p, f);
    if (c.nempty > 0 && !beta) {
        ZeroArgs za;
        za.rows = c.d_empty; za.nrows_empty = c.nempty; za.Z0 = Z0; za.Z1 = Z1; za.zsplit = zsplit; za.f = f;
        const long long total = (long long)c.nempty * (f / t.vw);
        const unsigned grid = (unsigned)((total + 255) / 256);
        if (t.vw == 4) zero_rows_kernel<4><<<grid, 256, 0, st>>>(za);
        else zero_rows_kernel<1><<<grid, 256, 0, st>>>(za);
        ++p->launches;
    }
    SpmmArgs a;
    a.blocks = sc.d_blocks; a.nblocks = sc.nblocks;
    a.pieces = c.d_cw;
    a.H0 = H0; a.H1 = H1; a.split = split;
    a.Z0 = Z0; a.Z1 = Z1; a.zsplit = zsplit;
    a.rowids = c.d_rowids;
    a.partial = sc.d_partial; a.f = f; a.beta = beta;
    a.relu = relu; a.final = (relu && use_final) ? c.d_final : nullptr;
    if (sc.nblocks > 0 && ring) {
        const int vpl = (f % 256 == 0) ? 2 : 1;
        const int tiles = f / (128 * vpl);
        const bool halo = (H1 != nullptr);
        int mode = p->opt_kernel == 6 ? 1 : (p->opt_kernel == 5 ? 0 : 2);   // default: tile::gather4
        // ring shape from the options: ring_slots = 16 | 32 | 64, ring_groups = 2 | 4 (only with 64 slots)
        const int64_t want_slots = c.tuned_slots > 0 ? c.tuned_slots : p->opt_ring_slots;
        int shape = want_slots <= 16 ? 0 : (want_slots <= 32 ? 1 : (p->opt_ring_groups == 4 ? 3 : 2));
        CUtensorMap tm0, tm1;
        if (mode == 2) {
            // H0 holds the columns below `split` (all of them when there is no halo slab), H1 the rest
            const int tile = 128 * vpl;
            bool ok = make_row_map(&tm0, H0, 1 << 30, f, tile, 1);
            if (ok && H1) ok = make_row_map(&tm1, H1, 1 << 30, f, tile, 1);
            else if (ok) tm1 = tm0;
            if (!ok) mode = 0;                                   // no driver entry point: 1-D bulk copies
        }
        if (mode == 1) shape = 0;
        if (mode == 0 && shape > 1) shape = 1;
        const int g = kRingShapes[shape].g, ng = kRingShapes[shape].ng;
        ring_fn fn = mode == 2 ? nullptr : pick_ring(vpl, shape, mode, halo);
        ring_g4_fn fn4 = mode == 2 ? pick_ring_g4(vpl, shape, halo) : nullptr;
        const void* fptr = mode == 2 ? (const void*)fn4 : (const void*)fn;
        const size_t smem = ring_smem_bytes(vpl, g * ng, ng);
        // opt-in to > 48 KB of dynamic shared memory, once per kernel instance
        const int slot = (((vpl - 1) * 4 + shape) * 3 + mode) * 2 + (halo ? 1 : 0);
        if (!p->ring_attr_set[slot]) {
            CU(p, cudaFuncSetAttribute(fptr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int nb = 0;
            CU(p, cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(&nb, fptr, kRingWarps * 32, smem, 0));
            p->ring_ctas_per_sm[slot] = std::max(nb, 1);
            p->ring_attr_set[slot] = true;
        }
        RingArgs ra;
        ra.counter = nullptr; ra.hub = nullptr; ra.nhub = 0;
        dim3 grid((unsigned)((sc.nblocks + kRingWarps - 1) / kRingWarps), (unsigned)tiles);
        // Persistent CTAs own their SM (shared memory + registers) until the whole launch is done; a put / NCCL kernel
        // of the exchange stream would then wait behind the SpMM it is supposed to overlap (measured at 8 GPUs: step =
        // sum of puts + sum of SpMMs). Multi-rank plans with overlap therefore run one block per warp (CTAs retire
        // every few dozen microseconds and the higher-priority exchange kernels take the freed slots) unless
        // `persistent_multi` asks otherwise.
        const bool persistent = p->opt_persistent && (p->k == 1 || !p->opt_overlap || p->opt_persistent_multi);
        if (persistent) {
            CU(p, cudaMemsetAsync(p->d_counter, 0, 64 * sizeof(unsigned int), st));
            ra.counter = p->d_counter;
            grid.x = std::min<unsigned>(grid.x, (unsigned)(p->num_sms * p->ring_ctas_per_sm[slot]));
        }
        if (mode == 2) fn4<<<grid, kRingWarps * 32, smem, st>>>(a, ra, tm0, tm1);
        else fn<<<grid, kRingWarps * 32, smem, st>>>(a, ra);
        ++p->launches;
    } else if (sc.nblocks > 0) {
        const int groups_per_cta = kSpmmThreads / t.lpe;
        dim3 grid((unsigned)((sc.nblocks + groups_per_cta - 1) / groups_per_cta), (unsigned)t.tiles);
        const bool halo = (H1 != nullptr);
        spmm_fn fn = (t.vw == 4) ? pick_lpe<4>(t.lpe, t.vpl, halo) : pick_lpe<1>(t.lpe, t.vpl, halo);
        fn<<<grid, kSpmmThreads, 0, st>>>(a);
        ++p->launches;
    }
    if (sc.nlong > 0) {
        FixupArgs fa;
        fa.long_rows = sc.d_long; fa.nlong = sc.nlong; fa.partial = sc.d_partial;
        fa.Z0 = Z0; fa.Z1 = Z1; fa.zsplit = zsplit; fa.rowids = c.d_rowids; fa.f = f; fa.beta = beta;
        fa.relu = a.relu; fa.final = a.final;
        const int nvec = f / t.vw;
        const unsigned grid = (unsigned)sc.nlong * (unsigned)((nvec + 31) / 32);
        if (t.vw == 4) spmm_fixup_kernel<4><<<grid, 32 * kFixupGroups, 0, st>>>(fa);
        else spmm_fixup_kernel<1><<<grid, 32 * kFixupGroups, 0, st>>>(fa);
        ++p->launches;
    }
    CU(p, cudaGetLastError());
    return 0;
}

int check_f(pgcn_plan* p, int f)
{
    if (!p) return fail(nullptr, PGCN_ERR_INVALID, "null plan");
    if (f <= 0 || f > p->f_max) return fail(p, PGCN_ERR_INVALID, "f=%d outside (0, f_max=%d]", f, p->f_max);
    return 0;
}

unsigned grid_for(long long total)
{
    long long g = (total + 255) / 256;
    return (unsigned)std::max<long long>(1, std::min<long long>(g, 148LL * 32));
}

int launch_pack(pgcn_plan* p, const float* H, float* slab, int f, cudaStream_t st)
{
    if (p->S == 0) return 0;
    PackArgs a;
    a.send_idx = p->d_send_idx; a.S = p->S; a.H = H; a.slab = slab; a.f = f;
    const int vw = (f % 4 == 0) ? 4 : 1;
    const unsigned grid = grid_for(p->S * (f / vw));
    if (vw == 4) pack_rows_kernel<4><<<grid, 256, 0, st>>>(a);
    else pack_rows_kernel<1><<<grid, 256, 0, st>>>(a);
    ++p->launches;
    CU(p, cudaGetLastError());
    return 0;
}

int launch_unpack(pgcn_plan* p, const float* recv, float* G, int f, cudaStream_t st)
{
    if (p->nb == 0) return 0;
    UnpackArgs a;
    a.brow = p->d_brow; a.bptr = p->d_bptr; a.bpos = p->d_bpos; a.nb = p->nb;
    a.recv = recv; a.G = G; a.f = f;
    const int vw = (f % 4 == 0) ? 4 : 1;
    const long long total = (long long)p->nb * (f / vw);
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (vw == 4) unpack_add_kernel<4><<<grid, 256, 0, st>>>(a);
    else unpack_add_kernel<1><<<grid, 256, 0, st>>>(a);
    ++p->launches;
    CU(p, cudaGetLastError());
    return 0;
}

// ---- transports ----------------------------------------------------------------------------

int nccl_exchange(pgcn_plan* p, const float* send, float* recv, int f, int reverse, cudaStream_t st)
{
    if (p->k == 1) return 0;
    if (!p->comm) return fail(p, PGCN_ERR_STATE, "exchange needs pgcn_comm_init (k=%d)", p->k);
    const std::vector<int64_t>& so = reverse ? p->recv_off : p->send_off;
    const std::vector<int64_t>& ro = reverse ? p->send_off : p->recv_off;
    NC(p, g_nccl.GroupStart());
    for (int q = 0; q < p->k; ++q) {
        if (q == p->rank) continue;
        const int64_t ns = so[q + 1] - so[q], nr = ro[q + 1] - ro[q];
        if (ns > 0) NC(p, g_nccl.Send(send + (size_t)so[q] * f, (size_t)ns * f, ncclFloat_, q, p->comm, st));
        if (nr > 0) NC(p, g_nccl.Recv(recv + (size_t)ro[q] * f, (size_t)nr * f, ncclFloat_, q, p->comm, st));
    }
    NC(p, g_nccl.GroupEnd());
    return 0;
}

float* arena_ptr(void* base, int64_t off) { return reinterpret_cast<float*>(static_cast<char*>(base) + off); }

unsigned long long* flag_slot(void* arena, int64_t off_flags, int slot)
{
    return reinterpret_cast<unsigned long long*>(static_cast<char*>(arena) + off_flags) + slot;
}

// Fused put of the rows bound for peer `dst` + epoch signal (peer-memory transport). `reverse`: halo partials of
// A^T g back to their owner (rows already in wire order), else boundary rows of H gathered through send_idx.
int p2p_put(pgcn_plan* p, int dst, const float* src, int f, bool reverse, int par, cudaStream_t st)
{
    PutArgs a;
    const P2PBlob& pb = p->peer_blob[dst];
    if (!reverse) {
        a.send_idx = p->d_send_idx; a.j0 = p->send_off[dst]; a.nrows = p->send_off[dst + 1] - p->send_off[dst];
        a.dst = arena_ptr(p->peer_arena[dst], pb.off_fwd[par]) + (size_t)pb.recv_off[p->rank] * f;
    } else {
        a.send_idx = nullptr; a.j0 = p->recv_off[dst]; a.nrows = p->recv_off[dst + 1] - p->recv_off[dst];
        a.dst = arena_ptr(p->peer_arena[dst], pb.off_bwd[par]) + (size_t)pb.send_off[p->rank] * f;
    }
    a.src = src; a.f = f;
    a.done = p->d_done + (reverse ? 32 : 0) + dst;     // kMaxPeers <= 16 destinations per direction
    a.flag = flag_slot(p->peer_arena[dst], pb.off_flags, p->rank);
    a.epoch = p->epoch;
    // enough CTAs to keep the NVLink store queues full, few enough not to crowd out the SpMM running beside it
    const long long items = a.nrows * (f / 4);
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((items + 255) / 256, 4LL * p->num_sms));
    put_rows_kernel<4><<<grid, 256, 0, st>>>(a);
    ++p->launches;
    CU(p, cudaGetLastError());
    return 0;
}

int p2p_wait(pgcn_plan* p, int src, cudaStream_t st)
{
    p2p_wait_kernel<<<1, 32, 0, st>>>(flag_slot(p->arena, p->off_flags, src), p->epoch);
    ++p->launches;
    CU(p, cudaGetLastError());
    return 0;
}

}  // namespace

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

const char* pgcn_version(void) { return "pgcn_b200 0.1 (sm_100a, CSR row-block SpMM + halo exchange)"; }

int pgcn_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, PGCN_ERR_NOGPU, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
    return n;
}

const char* pgcn_last_error(const pgcn_plan* plan) { return plan ? plan->err.c_str() : g_lib_error.c_str(); }

int pgcn_plan_create(const int32_t* rowptr, const int32_t* colidx, const float* vals,
                     int32_t m, int32_t h,
                     const int32_t* t_rowptr, const int32_t* t_colidx, const float* t_vals,
                     const int32_t* send_idx, const int64_t* send_off, const int64_t* recv_off,
                     int32_t k, int32_t rank, int32_t f_max, pgcn_plan** out)
{
    if (!out) return fail(nullptr, PGCN_ERR_INVALID, "out is null");
    *out = nullptr;
    if (!rowptr || !t_rowptr || !send_off || !recv_off) return fail(nullptr, PGCN_ERR_INVALID, "null index array");
    if ((int64_t)m + h >= (int64_t)kColMask) return fail(nullptr, PGCN_ERR_INVALID, "m + h must be below 2^30");
    if (m < 0 || h < 0 || k < 1 || rank < 0 || rank >= k || f_max < 1)
        return fail(nullptr, PGCN_ERR_INVALID, "bad sizes m=%d h=%d k=%d rank=%d f_max=%d", m, h, k, rank, f_max);
    if (recv_off[k] != h) return fail(nullptr, PGCN_ERR_INVALID, "recv_off[k]=%lld != h=%d", (long long)recv_off[k], h);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, PGCN_ERR_NOGPU, "no CUDA device: the PGCN B200 path has no CPU fallback");
    }
    const int64_t nnz = rowptr[m];
    if (nnz > 0 && (!colidx || !vals || !t_colidx || !t_vals)) return fail(nullptr, PGCN_ERR_INVALID, "null colidx/vals");
    if (t_rowptr[m + h] != nnz) return fail(nullptr, PGCN_ERR_INVALID, "transpose nnz mismatch");
    for (int64_t e = 0; e < nnz; ++e) {
        if (colidx[e] < 0 || colidx[e] >= m + h) return fail(nullptr, PGCN_ERR_INVALID, "colidx[%lld]=%d out of [0,%d)", (long long)e, colidx[e], m + h);
        if (t_colidx[e] < 0 || t_colidx[e] >= m) return fail(nullptr, PGCN_ERR_INVALID, "t_colidx[%lld]=%d out of [0,%d)", (long long)e, t_colidx[e], m);
    }
    const int64_t S = send_off[k];
    for (int64_t j = 0; j < S; ++j)
        if (!send_idx || send_idx[j] < 0 || send_idx[j] >= m) return fail(nullptr, PGCN_ERR_INVALID, "send_idx[%lld] out of range", (long long)j);

    pgcn_plan* p = new pgcn_plan();
    p->m = m; p->h = h; p->k = k; p->rank = rank; p->f_max = f_max; p->S = S;
    p->send_off.assign(send_off, send_off + k + 1);
    p->recv_off.assign(recv_off, recv_off + k + 1);
    cudaGetDevice(&p->device);

#define TRY(expr) do { int rc__ = (expr); if (rc__) { g_lib_error = p->err; pgcn_plan_destroy(p); return rc__; } } while (0)
    // column reference counts: forward columns = rows of the transpose and vice versa
    std::vector<int> refs_fwd((size_t)m + h), refs_tr((size_t)m);
    for (int r = 0; r < m + h; ++r) refs_fwd[r] = t_rowptr[r + 1] - t_rowptr[r];
    for (int r = 0; r < m; ++r) refs_tr[r] = rowptr[r + 1] - rowptr[r];
    // rows of H kept hot in L2: about half of the 126 MB L2, the rest is left to the streams
    if (const char* e = getenv("PGCN_HOT_MB")) p->opt_hot_mb = std::max<long long>(0, atoll(e));   // tuning knob
    const int64_t hot_rows = std::max<int64_t>(1, (p->opt_hot_mb << 20) / ((int64_t)f_max * 4));
    auto cold_threshold = [&](const std::vector<int>& refs) -> int {
        const int64_t ncols = (int64_t)refs.size();
        if (ncols <= hot_rows) return -1;                  // everything fits: nothing is cold
        std::vector<int> sorted(refs);
        std::nth_element(sorted.begin(), sorted.begin() + (ncols - hot_rows), sorted.end());
        return sorted[ncols - hot_rows];
    };
    const int cold_fwd = cold_threshold(refs_fwd), cold_tr = cold_threshold(refs_tr);
    TRY(csr_upload(p, p->fwd, m, rowptr, colidx, vals, nullptr, refs_fwd.data(), cold_fwd));
    TRY(csr_upload(p, p->tr, m + h, t_rowptr, t_colidx, t_vals, nullptr, refs_tr.data(), cold_tr, true));

    // distinct referenced columns / transposed rows (for the roofline's compulsory bytes)
    for (int r = 0; r < m + h; ++r) if (t_rowptr[r + 1] > t_rowptr[r]) ++p->cols_ref;
    p->rows_ref_t = 0;
    for (int r = 0; r < m; ++r) if (rowptr[r + 1] > rowptr[r]) ++p->rows_ref_t;

    // split A_local = [A_own | A_halo(peer) ...] (Parallel-GCN/main.c:271 then :295 per received block) so that
    // A_own * H_own overlaps the exchange and each peer's block is accumulated as soon as it has landed
    if (h > 0 && k > 1) {
        std::vector<int> o_rp(m + 1, 0), o_ci; std::vector<float> o_v;
        o_ci.reserve(nnz); o_v.reserve(nnz);
        std::vector<std::vector<int>> q_rp((size_t)k, std::vector<int>(1, 0)), q_ci((size_t)k), q_map((size_t)k);
        std::vector<std::vector<float>> q_v((size_t)k);
        std::vector<int> col_peer((size_t)h);
        for (int q = 0; q < k; ++q)
            for (int64_t c = recv_off[q]; c < recv_off[q + 1]; ++c) col_peer[(size_t)c] = q;
        std::vector<char> touched((size_t)k, 0);
        for (int r = 0; r < m; ++r) {
            for (int e = rowptr[r]; e < rowptr[r + 1]; ++e) {
                if (colidx[e] < m) { o_ci.push_back(colidx[e]); o_v.push_back(vals[e]); }
                else {
                    const int q = col_peer[(size_t)(colidx[e] - m)];
                    q_ci[q].push_back(colidx[e] - m);                    // slab-relative column
                    q_v[q].push_back(vals[e]);
                    touched[q] = 1;
                }
            }
            o_rp[r + 1] = (int)o_ci.size();
            for (int q = 0; q < k; ++q)
                if (touched[q]) { q_map[q].push_back(r); q_rp[q].push_back((int)q_ci[q].size()); touched[q] = 0; }
        }
        TRY(csr_upload(p, p->own, m, o_rp.data(), o_ci.data(), o_v.data(), nullptr, refs_fwd.data(), cold_fwd, true));
        p->halo_q.resize((size_t)k);
        for (int q = 0; q < k; ++q) {
            if (q_map[q].empty()) continue;
            TRY(csr_upload(p, p->halo_q[q], (int)q_map[q].size(), q_rp[q].data(), q_ci[q].data(), q_v[q].data(), &q_map[q],
                           refs_fwd.data() + m, cold_fwd));
        }
        // fused ReLU epilogue: a row is clamped by the launch of the pipelined forward that writes it last — the
        // own-columns launch for rows without halo entries, else the last peer block (in this rank's step order)
        {
            std::vector<int> last_step((size_t)m, 0);
            for (int i = 1; i < k; ++i) {
                const int src = (rank - i + k) % k;
                for (int r : q_map[src]) last_step[(size_t)r] = i;
            }
            std::vector<unsigned char> fin;
            fin.resize((size_t)std::max(p->own.nrows_c, 1));
            for (int c = 0; c < p->own.nrows_c; ++c) {
                const int r = p->own.d_rowids ? p->own.h_rowids[(size_t)c] : c;
                fin[(size_t)c] = last_step[(size_t)r] == 0;
            }
            TRY(upload(p, &p->own.d_final, fin.data(), fin.size()));
            for (int i = 1; i < k; ++i) {
                const int src = (rank - i + k) % k;
                if (q_map[src].empty()) continue;
                fin.resize(q_map[src].size());
                for (size_t j = 0; j < q_map[src].size(); ++j) fin[j] = last_step[(size_t)q_map[src][j]] == i;
                TRY(upload(p, &p->halo_q[(size_t)src].d_final, fin.data(), fin.size()));
            }
        }
        csr_view(p->tr, p->tr_own, 0, m);
        p->tr_halo_q.resize((size_t)k);
        for (int q = 0; q < k; ++q)
            if (recv_off[q + 1] > recv_off[q]) csr_view(p->tr, p->tr_halo_q[q], m + (int)recv_off[q], m + (int)recv_off[q + 1]);
        p->have_split = true;
    }

    {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, p->device) == cudaSuccess) p->num_sms = prop.multiProcessorCount;
        std::vector<unsigned int> zeros(64, 0u);
        TRY(upload(p, &p->d_counter, zeros.data(), zeros.size()));
        TRY(upload(p, &p->d_done, zeros.data(), zeros.size()));
    }
    TRY(upload(p, &p->d_send_idx, send_idx, (size_t)S));
    // boundary CSR: for every owned row that appears in some send list, the slab positions
    {
        std::vector<std::pair<int, int>> pr((size_t)S);
        for (int64_t j = 0; j < S; ++j) pr[j] = std::make_pair(send_idx[j], (int)j);
        std::stable_sort(pr.begin(), pr.end(), [](const std::pair<int,int>& a, const std::pair<int,int>& b) { return a.first < b.first; });
        std::vector<int> brow, bptr(1, 0), bpos((size_t)S);
        for (int64_t j = 0; j < S; ++j) {
            if (j == 0 || pr[j].first != pr[j - 1].first) { if (j) bptr.push_back((int)j); brow.push_back(pr[j].first); }
            bpos[j] = pr[j].second;
        }
        if (S) bptr.push_back((int)S);
        p->nb = (int)brow.size();
        TRY(upload(p, &p->d_brow, brow.data(), brow.size()));
        TRY(upload(p, &p->d_bptr, bptr.data(), bptr.size()));
        TRY(upload(p, &p->d_bpos, bpos.data(), bpos.size()));
    }
    auto slab = [&](float** d, int64_t rows) -> int {
        CU(p, cudaMalloc((void**)d, std::max<size_t>((size_t)rows * f_max, 1) * sizeof(float)));
        return 0;
    };
    TRY(slab(&p->d_send_slab, S));
    TRY(slab(&p->d_halo_slab, h));
    TRY(slab(&p->d_rrecv_slab, S));
    TRY(slab(&p->d_hsend_slab, h));
    {
        // the exchange stream outranks the compute stream: its (small) kernels must get SM slots while an SpMM
        // grid is draining, because a neighbour is waiting for their stores
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        cudaError_t e1 = cudaStreamCreateWithPriority(&p->comm_stream, cudaStreamNonBlocking, prio_hi);
        if (e1 == cudaSuccess) e1 = cudaStreamCreateWithFlags(&p->host_stream, cudaStreamNonBlocking);
        cudaError_t e2 = cudaEventCreateWithFlags(&p->ev_a, cudaEventDisableTiming);
        cudaError_t e3 = cudaEventCreateWithFlags(&p->ev_b, cudaEventDisableTiming);
        p->ev_step.assign((size_t)k, nullptr);
        for (int q = 0; q < k && e3 == cudaSuccess; ++q) e3 = cudaEventCreateWithFlags(&p->ev_step[(size_t)q], cudaEventDisableTiming);
        if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
            fail(p, PGCN_ERR_CUDA, "stream/event creation failed");
            g_lib_error = p->err; pgcn_plan_destroy(p); return PGCN_ERR_CUDA;
        }
    }
#undef TRY
    *out = p;
    return 0;
}

int pgcn_plan_destroy(pgcn_plan* p)
{
    if (!p) return 0;
    cudaSetDevice(p->device);
    cudaDeviceSynchronize();
    if (p->comm && g_nccl.ok && !p->comm_borrowed) g_nccl.CommDestroy(p->comm);
    for (int q = 0; q < kMaxPeers; ++q)
        if (p->peer_arena[q] && q != p->rank && !p->peer_local[q]) cudaIpcCloseMemHandle(p->peer_arena[q]);
    cudaFree(p->arena);
    csr_free(p->tr_own);
    for (auto& c : p->tr_halo_q) csr_free(c);
    for (auto& c : p->halo_q) csr_free(c);
    csr_free(p->fwd); csr_free(p->tr); csr_free(p->own);
    for (cudaEvent_t e : p->ev_step) if (e) cudaEventDestroy(e);
    cudaFree(p->d_done);
    cudaFree(p->d_send_idx);
    cudaFree(p->d_brow); cudaFree(p->d_bptr); cudaFree(p->d_bpos);
    cudaFree(p->d_send_slab); cudaFree(p->d_halo_slab); cudaFree(p->d_rrecv_slab); cudaFree(p->d_hsend_slab);
    for (int i = 0; i < 2; ++i) {
        cudaFree(p->d_hostH[i]); cudaFree(p->d_hostZ[i]);
        if (p->ev_in[i]) cudaEventDestroy(p->ev_in[i]);
        if (p->ev_comp[i]) cudaEventDestroy(p->ev_comp[i]);
        if (p->ev_out[i]) cudaEventDestroy(p->ev_out[i]);
    }
    if (p->s_in) cudaStreamDestroy(p->s_in);
    if (p->s_out) cudaStreamDestroy(p->s_out);
    cudaFree(p->d_counter);
    if (p->comm_stream) cudaStreamDestroy(p->comm_stream);
    if (p->host_stream) cudaStreamDestroy(p->host_stream);
    if (p->ev_a) cudaEventDestroy(p->ev_a);
    if (p->ev_b) cudaEventDestroy(p->ev_b);
    cudaGetLastError();
    delete p;
    return 0;
}

int pgcn_plan_set_option(pgcn_plan* p, const char* name, int64_t value)
{
    if (!p || !name) return fail(p, PGCN_ERR_INVALID, "null argument");
    const std::string n(name);
    auto clear_tuned = [&]() {
        DevCsr* all[] = {&p->fwd, &p->tr, &p->own, &p->tr_own};
        for (DevCsr* c : all) { c->tuned_epb[0] = c->tuned_epb[1] = 0; c->tuned_slots = 0; }
        for (auto& c : p->halo_q) { c.tuned_epb[0] = c.tuned_epb[1] = 0; c.tuned_slots = 0; }
        for (auto& c : p->tr_halo_q) { c.tuned_epb[0] = c.tuned_epb[1] = 0; c.tuned_slots = 0; }
    };
    if (n == "edges_per_block" || n == "ring_edges_per_block" || n == "ring_slots") clear_tuned();   // explicit beats tuned
    if (n == "edges_per_block") p->opt_epb = value;
    else if (n == "kernel") p->opt_kernel = value;
    else if (n == "ring_slots") p->opt_ring_slots = value;
    else if (n == "ring_edges_per_block") p->opt_ring_epb = value;
    else if (n == "ring_long_row") p->opt_ring_long = value;
    else if (n == "persistent") p->opt_persistent = value;
    else if (n == "persistent_multi") p->opt_persistent_multi = value;
    else if (n == "ring_groups") p->opt_ring_groups = value;
    else if (n == "long_row") p->opt_long = value;
    else if (n == "tile_floats") p->opt_tile = value;
    else if (n == "hot_mb")
        // the cold-column marks are baked into the pair arrays at pgcn_plan_create: a later change would be a
        // silent no-op, so it is refused (use the PGCN_HOT_MB environment variable before creating the plan)
        return fail(p, PGCN_ERR_STATE, "hot_mb is fixed at plan creation (set PGCN_HOT_MB before pgcn_plan_create)");
    else if (n == "overlap") p->opt_overlap = value;
    else if (n == "relu") p->opt_relu = value ? 1 : 0;
    else if (n == "p2p") {
        // 0 = never use the peer transport even though pgcn_p2p_import succeeded here (another rank could not map
        // its peers: every rank must then fall back to NCCL together); 1 re-enables it when the arenas are mapped.
        p->opt_p2p = value ? 1 : 0;
        p->p2p = p->opt_p2p && p->arena && p->peer_arena[p->rank];
    }
    else return fail(p, PGCN_ERR_INVALID, "unknown option '%s'", name);
    return 0;
}

int64_t pgcn_plan_get_option(const pgcn_plan* p, const char* name)
{
    if (!p || !name) return PGCN_ERR_INVALID;
    const std::string n(name);
    if (n == "edges_per_block") return p->fwd.tuned_epb[0] > 0 ? p->fwd.tuned_epb[0] : p->opt_epb;
    if (n == "kernel") return p->opt_kernel;
    if (n == "ring_slots") return p->fwd.tuned_slots > 0 ? p->fwd.tuned_slots : p->opt_ring_slots;
    if (n == "ring_edges_per_block") return p->fwd.tuned_epb[1] > 0 ? p->fwd.tuned_epb[1] : p->opt_ring_epb;
    if (n == "ring_long_row") return p->opt_ring_long;
    if (n == "persistent") return p->opt_persistent;
    if (n == "persistent_multi") return p->opt_persistent_multi;
    if (n == "ring_groups") return p->opt_ring_groups;
    if (n == "long_row") return p->opt_long;
    if (n == "tile_floats") return p->opt_tile;
    if (n == "hot_mb") return p->opt_hot_mb;
    if (n == "overlap") return p->opt_overlap;
    if (n == "relu") return p->opt_relu;
    if (n == "p2p") return p->p2p ? 1 : 0;
    if (n == "nccl") return p->comm ? 1 : 0;
    if (n == "blocks_fwd") return p->fwd.sched[0].nblocks;
    if (n == "long_rows_fwd") return p->fwd.sched[0].nlong;
    if (n == "ring_blocks_fwd") return p->fwd.sched[1].nblocks;
    if (n == "ring_long_rows_fwd") return p->fwd.sched[1].nlong;
    return PGCN_ERR_INVALID;
}

int64_t pgcn_debug_schedule(const int32_t* rowptr, int32_t nrows, int64_t edges_per_block, int64_t long_row,
                            int32_t* blocks_out, int64_t cap_blocks, int32_t* nlong_out, int32_t* nslots_out)
{
    if (!rowptr || nrows < 0 || edges_per_block < 8) return fail(nullptr, PGCN_ERR_INVALID, "bad schedule arguments");
    std::vector<int4> blocks, longs;
    int nslots = 0;
    make_schedule(rowptr, nrows, edges_per_block, long_row > 0 ? long_row : 4 * edges_per_block, blocks, longs, nslots);
    if (nlong_out) *nlong_out = (int32_t)longs.size();
    if (nslots_out) *nslots_out = nslots;
    if (blocks_out) {
        const int64_t n = std::min<int64_t>((int64_t)blocks.size(), cap_blocks);
        for (int64_t i = 0; i < n; ++i) {
            blocks_out[4 * i] = blocks[i].x; blocks_out[4 * i + 1] = blocks[i].y;
            blocks_out[4 * i + 2] = blocks[i].z; blocks_out[4 * i + 3] = blocks[i].w;
        }
    }
    return (int64_t)blocks.size();
}

// Autotune: every matrix of the plan (forward, transposed, and — for k > 1 — the own-column part, each peer's halo
// block and the transposed row ranges the pipelined backward launches) gets its own schedule parameters, timed on
// zero-filled scratch operands of the right shapes. Returns the edges-per-block chosen for the forward matrix.
int pgcn_plan_autotune(pgcn_plan* p, int32_t f)
{
    int rc = check_f(p, f);
    if (rc) return rc;
    if (p->fwd.nnz == 0) return (int)p->opt_epb;
    CU(p, cudaSetDevice(p->device));
    float *H0 = nullptr, *H1 = nullptr, *Z0 = nullptr, *Z1 = nullptr;
    const size_t bm = std::max<size_t>((size_t)p->m * f, 1) * 4, bh = std::max<size_t>((size_t)p->h * f, 1) * 4;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    auto cleanup = [&]() {
        cudaFree(H0); cudaFree(H1); cudaFree(Z0); cudaFree(Z1);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    };
    if (cudaMalloc((void**)&H0, bm) != cudaSuccess || cudaMalloc((void**)&H1, bh) != cudaSuccess ||
        cudaMalloc((void**)&Z0, bm) != cudaSuccess || cudaMalloc((void**)&Z1, bh) != cudaSuccess) {
        cleanup(); cudaGetLastError();
        return fail(p, PGCN_ERR_CUDA, "autotune: scratch allocation failed");
    }
    cudaStream_t st = p->host_stream;
    cudaMemsetAsync(H0, 0, bm, st); cudaMemsetAsync(H1, 0, bh, st);
    cudaMemsetAsync(Z0, 0, bm, st); cudaMemsetAsync(Z1, 0, bh, st);
    cudaEventCreate(&e0); cudaEventCreate(&e1);

    struct Cand { int64_t epb; int slots; };
    static const Cand ring_cand[] = {{256, 16}, {512, 16}, {1024, 16}, {512, 32}, {1024, 32}};
    static const Cand reg_cand[] = {{96, 0}, {128, 0}, {144, 0}, {160, 0}, {192, 0}, {256, 0}};
    // one matrix: h0/h1/split = gathered operand(s), z0/z1/zsplit = outputs
    auto tune = [&](DevCsr& c, const float* h0, const float* h1, int split, float* z0, float* z1, int zsplit, int beta) -> int {
        if (c.nnz == 0 || c.nrows == 0) return 0;
        const bool ring = use_ring(p, h0, h1, f);
        const Cand* cand = ring ? ring_cand : reg_cand;
        const int ncand = ring ? (int)(sizeof ring_cand / sizeof ring_cand[0]) : (int)(sizeof reg_cand / sizeof reg_cand[0]);
        const int which = ring ? 1 : 0;
        const int64_t keep_epb = c.tuned_epb[which];
        const int keep_slots = c.tuned_slots;
        int best = -1;
        float best_ms = 1e30f;
        for (int i = 0; i < ncand; ++i) {
            c.tuned_epb[which] = cand[i].epb;
            if (ring) c.tuned_slots = cand[i].slots;
            float ms_min = 1e30f;
            for (int it = 0; it < 4; ++it) {                 // first pass also builds the schedule
                cudaEventRecord(e0, st);
                int r2 = launch_spmm(p, c, h0, h1, split, z0, z1, zsplit, f, beta, st);
                cudaEventRecord(e1, st);
                if (r2 || cudaEventSynchronize(e1) != cudaSuccess) {
                    c.tuned_epb[which] = keep_epb; c.tuned_slots = keep_slots;
                    return r2 ? r2 : fail(p, PGCN_ERR_CUDA, "autotune: kernel failed");
                }
                float ms = 0.f;
                cudaEventElapsedTime(&ms, e0, e1);
                if (it > 0) ms_min = std::min(ms_min, ms);
            }
            if (ms_min < best_ms) { best_ms = ms_min; best = i; }
        }
        c.tuned_epb[which] = cand[best].epb;
        if (ring) c.tuned_slots = cand[best].slots;
        return 0;
    };
#define TUNE(...) do { if ((rc = tune(__VA_ARGS__))) { cleanup(); return rc; } } while (0)
    TUNE(p->fwd, H0, p->h ? H1 : nullptr, p->m, Z0, nullptr, p->m, 0);
    TUNE(p->tr, H0, nullptr, p->m, Z0, Z1, p->m, 0);
    if (p->have_split) {
        TUNE(p->own, H0, nullptr, p->m, Z0, nullptr, p->m, 0);
        TUNE(p->tr_own, H0, nullptr, p->m, Z0, Z1, p->m, 0);
        for (int q = 0; q < p->k; ++q) {
            TUNE(p->halo_q[(size_t)q], H1, nullptr, p->h, Z0, nullptr, p->m, 1);
            TUNE(p->tr_halo_q[(size_t)q], H0, nullptr, p->m, Z0, Z1, p->m, 0);
        }
    }
#undef TUNE
    cleanup();
    const int which = use_ring(p, H0, nullptr, f) ? 1 : 0;     // H0 was cudaMalloc'ed: aligned
    return (int)(p->fwd.tuned_epb[which] > 0 ? p->fwd.tuned_epb[which] : (which ? p->opt_ring_epb : p->opt_epb));
}

void* pgcn_plan_slab(pgcn_plan* p, int which)
{
    if (!p) return nullptr;
    switch (which) {
        case 0: return p->d_send_slab;
        case 1: return p->d_halo_slab;
        case 2: return p->d_rrecv_slab;
        case 3: return p->d_hsend_slab;
        default: return nullptr;
    }
}

int pgcn_algorithmic_bytes(const pgcn_plan* p, int32_t f, pgcn_bytes* o)
{
    if (!p || !o) return PGCN_ERR_INVALID;
    const int64_t nnz = p->fwd.nnz, m = p->m, h = p->h, S = p->S, F = f;
    o->nnz = nnz; o->m = m; o->h = h; o->cols_ref = p->cols_ref;
    o->spmm_fwd = 8 * nnz + 4 * (m + 1) + 4 * F * p->cols_ref + 4 * F * m;
    o->spmm_bwd = 8 * nnz + 4 * (m + h + 1) + 4 * F * p->rows_ref_t + 4 * F * (m + h);
    o->gather_fwd = nnz * (8 + 4 * F) + 4 * (m + 1) + 4 * F * m;
    o->xchg_out = 4 * F * S;
    o->xchg_in = 4 * F * h;
    o->pack = 2 * 4 * F * S;
    return 0;
}

int64_t pgcn_launch_count(const pgcn_plan* p) { return p ? p->launches : 0; }

// ---- communicator --------------------------------------------------------------------------

int pgcn_comm_unique_id(void* id128)
{
    if (!id128) return fail(nullptr, PGCN_ERR_INVALID, "null id buffer");
    if (!load_nccl()) return fail(nullptr, PGCN_ERR_NCCL, "libnccl.so.2 not found in this process");
    ncclUniqueId id;
    NC(nullptr, g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return 0;
}

int pgcn_comm_init(pgcn_plan* p, const void* id128)
{
    if (!p || !id128) return fail(p, PGCN_ERR_INVALID, "null argument");
    if (!load_nccl()) return fail(p, PGCN_ERR_NCCL, "libnccl.so.2 not found in this process");
    if (p->comm) return 0;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    CU(p, cudaSetDevice(p->device));
    NC(p, g_nccl.CommInitRank(&p->comm, p->k, id, p->rank));
    return 0;
}

int pgcn_comm_share(pgcn_plan* p, pgcn_plan* owner)
{
    if (!p || !owner) return fail(p, PGCN_ERR_INVALID, "null argument");
    if (!owner->comm) return fail(p, PGCN_ERR_STATE, "the owner plan has no communicator (pgcn_comm_init first)");
    if (p->k != owner->k || p->rank != owner->rank || p->device != owner->device)
        return fail(p, PGCN_ERR_INVALID, "plans of different rank / size / device cannot share a communicator");
    if (p->comm && !p->comm_borrowed) return fail(p, PGCN_ERR_STATE, "plan already owns a communicator");
    p->comm = owner->comm;
    p->comm_borrowed = true;
    return 0;
}

int pgcn_p2p_export(pgcn_plan* p, void* handle_out)
{
    if (!p || !handle_out) return fail(p, PGCN_ERR_INVALID, "null argument");
    if (p->k > kMaxPeers) return fail(p, PGCN_ERR_INVALID, "peer transport supports k <= %d", kMaxPeers);
    if (!p->arena) {
        auto align = [](int64_t x) { return (x + 255) / 256 * 256; };
        int64_t off = 0;
        p->off_flags = off; off = align(off + (int64_t)kMaxPeers * 8);
        const int64_t fwd_bytes = align((int64_t)p->h * p->f_max * 4), bwd_bytes = align((int64_t)p->S * p->f_max * 4);
        for (int i = 0; i < 2; ++i) { p->off_fwd[i] = off; off += std::max<int64_t>(fwd_bytes, 256); }
        for (int i = 0; i < 2; ++i) { p->off_bwd[i] = off; off += std::max<int64_t>(bwd_bytes, 256); }
        p->arena_bytes = off;
        CU(p, cudaMalloc(&p->arena, (size_t)off));
        CU(p, cudaMemset(p->arena, 0, (size_t)off));
        CU(p, cudaDeviceSynchronize());
    }
    P2PBlob b;
    memset(&b, 0, sizeof b);
    CU(p, cudaIpcGetMemHandle(&b.ipc, p->arena));
    b.arena_bytes = p->arena_bytes; b.off_flags = p->off_flags;
    for (int i = 0; i < 2; ++i) { b.off_fwd[i] = p->off_fwd[i]; b.off_bwd[i] = p->off_bwd[i]; }
    b.k = p->k; b.rank = p->rank; b.f_max = p->f_max;
    b.pid = (int64_t)getpid(); b.local_ptr = p->arena;
    for (int i = 0; i <= p->k; ++i) { b.send_off[i] = p->send_off[i]; b.recv_off[i] = p->recv_off[i]; }
    memset(handle_out, 0, PGCN_P2P_HANDLE_BYTES);
    memcpy(handle_out, &b, sizeof b);
    return 0;
}

int pgcn_p2p_import(pgcn_plan* p, const void* handles_k)
{
    if (!p || !handles_k) return fail(p, PGCN_ERR_INVALID, "null argument");
    if (!p->arena) return fail(p, PGCN_ERR_STATE, "call pgcn_p2p_export first");
    const char* base = static_cast<const char*>(handles_k);
    for (int q = 0; q < p->k; ++q) {
        memcpy(&p->peer_blob[q], base + (size_t)q * PGCN_P2P_HANDLE_BYTES, sizeof(P2PBlob));
        const P2PBlob& b = p->peer_blob[q];
        if (b.k != p->k || b.rank != q || b.f_max != p->f_max)
            return fail(p, PGCN_ERR_INVALID, "peer blob %d inconsistent (k=%d rank=%d f_max=%d)", q, b.k, b.rank, b.f_max);
        // wire-order invariant: what I send to q is what q expects from me (GPU/PGCN.py:47-48)
        if (b.recv_off[p->rank + 1] - b.recv_off[p->rank] != p->send_off[q + 1] - p->send_off[q])
            return fail(p, PGCN_ERR_INVALID, "send/recv count mismatch with peer %d", q);
        if (q == p->rank) { p->peer_arena[q] = p->arena; continue; }
        if (p->peer_arena[q]) continue;                     // already mapped (import called twice)
        if (b.pid == (int64_t)getpid()) {                   // a plan of this very process: plain device pointer
            p->peer_arena[q] = b.local_ptr;
            p->peer_local[q] = true;
            continue;
        }
        CU(p, cudaIpcOpenMemHandle(&p->peer_arena[q], b.ipc, cudaIpcMemLazyEnablePeerAccess));
    }
    preload_kernels();
    p->p2p = p->opt_p2p != 0;
    return 0;
}

// ---- hot path --------------------------------------------------------------------------------

int pgcn_spmm(pgcn_plan* p, int transpose, const float* H_own, const float* H_halo,
              float* Z, float* Z_halo, int32_t f, void* stream)
{
    int rc = check_f(p, f);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (transpose == 2 || transpose == 3) {
        // the two halves of the overlapped forward, individually callable (k > 1 plans only)
        if (!p->have_split) return fail(p, PGCN_ERR_STATE, "plan has no own/halo split (k == 1 or h == 0)");
        if (!Z) return fail(p, PGCN_ERR_INVALID, "null Z");
        if (transpose == 2) {
            if (!H_own) return fail(p, PGCN_ERR_INVALID, "null H_own");
            return launch_spmm(p, p->own, H_own, nullptr, p->m, Z, nullptr, p->m, f, 0, st);
        }
        if (!H_halo) return fail(p, PGCN_ERR_INVALID, "null H_halo");
        for (int q = 0; q < p->k; ++q) {                  // Z += A_halo(q) * H_halo, one source peer after the other
            if (p->halo_q[(size_t)q].nrows == 0) continue;
            int rc2 = launch_spmm(p, p->halo_q[(size_t)q], H_halo, nullptr, p->h, Z, nullptr, p->m, f, 1, st);
            if (rc2) return rc2;
        }
        return 0;
    }
    if (!transpose) {
        if (p->m > 0 && (!H_own || !Z)) return fail(p, PGCN_ERR_INVALID, "null H_own/Z");
        if (p->h > 0 && !H_halo) return fail(p, PGCN_ERR_INVALID, "h=%d but H_halo is null", p->h);
        return launch_spmm(p, p->fwd, H_own, p->h > 0 ? H_halo : nullptr, p->m, Z, nullptr, p->m, f, 0, st);
    }
    if (p->m > 0 && (!H_own || !Z)) return fail(p, PGCN_ERR_INVALID, "null gZ/G");
    if (p->h > 0 && !Z_halo) return fail(p, PGCN_ERR_INVALID, "h=%d but Z_halo is null", p->h);
    return launch_spmm(p, p->tr, H_own, nullptr, p->m, Z, Z_halo, p->m, f, 0, st);
}

int pgcn_pack(pgcn_plan* p, const float* H, float* send_slab, int32_t f, void* stream)
{
    int rc = check_f(p, f);
    if (rc) return rc;
    if (p->S > 0 && (!H || !send_slab)) return fail(p, PGCN_ERR_INVALID, "null H/send_slab");
    return launch_pack(p, H, send_slab, f, (cudaStream_t)stream);
}

int pgcn_exchange(pgcn_plan* p, const float* send_slab, float* recv_slab, int32_t f, int reverse, void* stream)
{
    int rc = check_f(p, f);
    if (rc) return rc;
    return nccl_exchange(p, send_slab, recv_slab, f, reverse, (cudaStream_t)stream);
}

int pgcn_unpack_add(pgcn_plan* p, const float* recv_slab, float* G_own, int32_t f, void* stream)
{
    int rc = check_f(p, f);
    if (rc) return rc;
    if (p->S > 0 && (!recv_slab || !G_own)) return fail(p, PGCN_ERR_INVALID, "null recv_slab/G_own");
    return launch_unpack(p, recv_slab, G_own, f, (cudaStream_t)stream);
}

// Exchange schedule shared by both transports and both directions: at step i = 1 .. k-1 a rank sends to
// (rank + i) % k and receives from (rank - i) % k — every pair is active exactly once per step, and a receiver
// sees its sources arrive one after the other, so each block can be consumed while the next is in flight.
static inline int step_dst(const pgcn_plan* p, int i) { return (p->rank + i) % p->k; }
static inline int step_src(const pgcn_plan* p, int i) { return (p->rank - i + p->k) % p->k; }

static int nccl_step(pgcn_plan* p, const float* send, float* recv, int f, int reverse, int i, cudaStream_t st)
{
    const std::vector<int64_t>& so = reverse ? p->recv_off : p->send_off;
    const std::vector<int64_t>& ro = reverse ? p->send_off : p->recv_off;
    const int d = step_dst(p, i), s = step_src(p, i);
    const int64_t ns = so[d + 1] - so[d], nr = ro[s + 1] - ro[s];
    if (ns == 0 && nr == 0) return 0;
    NC(p, g_nccl.GroupStart());
    if (ns > 0) NC(p, g_nccl.Send(send + (size_t)so[d] * f, (size_t)ns * f, ncclFloat_, d, p->comm, st));
    if (nr > 0) NC(p, g_nccl.Recv(recv + (size_t)ro[s] * f, (size_t)nr * f, ncclFloat_, s, p->comm, st));
    NC(p, g_nccl.GroupEnd());
    return 0;
}

int pgcn_forward(pgcn_plan* p, const float* H_own, float* Z, int32_t f, void* stream)
{
    int rc = check_f(p, f);
    if (rc) return rc;
    if (p->m > 0 && (!H_own || !Z)) return fail(p, PGCN_ERR_INVALID, "null H_own/Z");
    cudaStream_t st = (cudaStream_t)stream;
    const int relu = (int)p->opt_relu;
    if (p->k == 1)
        return launch_spmm(p, p->fwd, H_own, p->h > 0 ? p->d_halo_slab : nullptr, p->m, Z, nullptr, p->m, f, 0, st, relu);

    const bool use_p2p = p->p2p && (f % 4 == 0);
    if (!use_p2p && !p->comm) return fail(p, PGCN_ERR_STATE, "k=%d: call pgcn_comm_init or pgcn_p2p_import first", p->k);
    const bool split = p->have_split && p->opt_overlap;
    const int k = p->k;
    float* halo = p->d_halo_slab;
    int par = 0;
    if (use_p2p) {
        ++p->epoch;
        par = (int)(p->epoch & 1);
        halo = arena_ptr(p->arena, p->off_fwd[par]);
    }
    cudaStream_t cs = split ? p->comm_stream : st;
    if (split) {
        CU(p, cudaEventRecord(p->ev_a, st));
        CU(p, cudaStreamWaitEvent(cs, p->ev_a, 0));
    }
    // ---- send side (exchange stream): one peer after the other, in step order
    if (use_p2p) {
        for (int i = 1; i < k; ++i)
            if ((rc = p2p_put(p, step_dst(p, i), H_own, f, false, par, cs))) return rc;
        if (split) CU(p, cudaEventRecord(p->ev_b, cs));                   // H_own is free for the caller after this
    } else {
        if ((rc = launch_pack(p, H_own, p->d_send_slab, f, cs))) return rc;
        for (int i = 1; i < k; ++i) {
            if ((rc = nccl_step(p, p->d_send_slab, halo, f, 0, i, cs))) return rc;
            if (split) CU(p, cudaEventRecord(p->ev_step[(size_t)i], cs));
        }
    }
    if (!split) {
        // no overlap requested (or nothing to split): wait for every block, then one pass over [own | halo]
        if (use_p2p)
            for (int i = 1; i < k; ++i)
                if ((rc = p2p_wait(p, step_src(p, i), st))) return rc;
        return launch_spmm(p, p->fwd, H_own, p->h > 0 ? halo : nullptr, p->m, Z, nullptr, p->m, f, 0, st, relu);
    }
    // ---- compute side: own columns while the rows travel, then each source's block as soon as it has landed
    if ((rc = launch_spmm(p, p->own, H_own, nullptr, p->m, Z, nullptr, p->m, f, 0, st, relu, true))) return rc;
    for (int i = 1; i < k; ++i) {
        const int src = step_src(p, i);
        if (use_p2p) { if ((rc = p2p_wait(p, src, st))) return rc; }
        else CU(p, cudaStreamWaitEvent(st, p->ev_step[(size_t)i], 0));
        if (p->halo_q[(size_t)src].nrows == 0) continue;
        if ((rc = launch_spmm(p, p->halo_q[(size_t)src], halo, nullptr, p->h, Z, nullptr, p->m, f, 1, st, relu, true))) return rc;
    }
    if (use_p2p) CU(p, cudaStreamWaitEvent(st, p->ev_b, 0));
    return 0;
}

int pgcn_backward(pgcn_plan* p, const float* gZ, float* G_own, int32_t f, void* stream)
{
    int rc = check_f(p, f);
    if (rc) return rc;
    if (p->m > 0 && (!gZ || !G_own)) return fail(p, PGCN_ERR_INVALID, "null gZ/G_own");
    cudaStream_t st = (cudaStream_t)stream;
    // A^T g : rows [0,m) -> G_own, rows [m,m+h) -> halo partials, already in reverse wire order
    if (p->k == 1) return launch_spmm(p, p->tr, gZ, nullptr, p->m, G_own, p->d_hsend_slab, p->m, f, 0, st);

    const bool use_p2p = p->p2p && (f % 4 == 0);
    if (!use_p2p && !p->comm) return fail(p, PGCN_ERR_STATE, "k=%d: call pgcn_comm_init or pgcn_p2p_import first", p->k);
    const bool split = p->have_split && p->opt_overlap;
    const int k = p->k;
    float* rrecv = p->d_rrecv_slab;
    int par = 0;
    if (use_p2p) {
        ++p->epoch;
        par = (int)(p->epoch & 1);
        rrecv = arena_ptr(p->arena, p->off_bwd[par]);
    }
    if (!split) {
        if ((rc = launch_spmm(p, p->tr, gZ, nullptr, p->m, G_own, p->d_hsend_slab, p->m, f, 0, st))) return rc;
        for (int i = 1; i < k; ++i) {
            if (use_p2p) { if ((rc = p2p_put(p, step_dst(p, i), p->d_hsend_slab, f, true, par, st))) return rc; }
            else if ((rc = nccl_step(p, p->d_hsend_slab, rrecv, f, 1, i, st))) return rc;
        }
        if (use_p2p)
            for (int i = 1; i < k; ++i)
                if ((rc = p2p_wait(p, step_src(p, i), st))) return rc;
        return launch_unpack(p, rrecv, G_own, f, st);
    }
    // ---- pipelined: the partials owed to each peer are computed first (in step order) and leave on the exchange
    // stream while the next peer's rows, and finally the own rows of A^T g, are still being computed
    cudaStream_t cs = p->comm_stream;
    for (int i = 1; i < k; ++i) {
        const int dst = step_dst(p, i);
        if (p->tr_halo_q[(size_t)dst].nrows > 0)
            if ((rc = launch_spmm(p, p->tr_halo_q[(size_t)dst], gZ, nullptr, p->m, G_own, p->d_hsend_slab, p->m, f, 0, st))) return rc;
        CU(p, cudaEventRecord(p->ev_step[(size_t)i], st));
        CU(p, cudaStreamWaitEvent(cs, p->ev_step[(size_t)i], 0));
        if (use_p2p) { if ((rc = p2p_put(p, dst, p->d_hsend_slab, f, true, par, cs))) return rc; }
        else if ((rc = nccl_step(p, p->d_hsend_slab, rrecv, f, 1, i, cs))) return rc;
    }
    CU(p, cudaEventRecord(p->ev_b, cs));
    if ((rc = launch_spmm(p, p->tr_own, gZ, nullptr, p->m, G_own, p->d_hsend_slab, p->m, f, 0, st))) return rc;
    if (use_p2p) {
        for (int i = 1; i < k; ++i)
            if ((rc = p2p_wait(p, step_src(p, i), st))) return rc;
    }
    CU(p, cudaStreamWaitEvent(st, p->ev_b, 0));      // NCCL: all blocks received; p2p: the send slab is free again
    return launch_unpack(p, rrecv, G_own, f, st);
}

static int host_slots(pgcn_plan* p, int64_t need)
{
    if (!p->s_in) {
        CU(p, cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
        CU(p, cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CU(p, cudaEventCreateWithFlags(&p->ev_in[i], cudaEventDisableTiming));
            CU(p, cudaEventCreateWithFlags(&p->ev_comp[i], cudaEventDisableTiming));
            CU(p, cudaEventCreateWithFlags(&p->ev_out[i], cudaEventDisableTiming));
        }
    }
    if (need > p->host_cap) {
        CU(p, cudaDeviceSynchronize());
        for (int i = 0; i < 2; ++i) {
            cudaFree(p->d_hostH[i]); cudaFree(p->d_hostZ[i]);
            p->d_hostH[i] = p->d_hostZ[i] = nullptr;
        }
        p->host_cap = 0;
        for (int i = 0; i < 2; ++i) {
            CU(p, cudaMalloc((void**)&p->d_hostH[i], std::max<size_t>((size_t)need, 1) * 4));
            CU(p, cudaMalloc((void**)&p->d_hostZ[i], std::max<size_t>((size_t)need, 1) * 4));
        }
        p->host_cap = need;
        p->host_steps = 0;
    }
    return 0;
}

int pgcn_forward_host_async(pgcn_plan* p, const float* H_host, float* Z_host, int32_t f)
{
    int rc = check_f(p, f);
    if (rc) return rc;
    if (p->m > 0 && (!H_host || !Z_host)) return fail(p, PGCN_ERR_INVALID, "null host buffer");
    CU(p, cudaSetDevice(p->device));
    const int64_t need = (int64_t)p->m * f;
    if ((rc = host_slots(p, need))) return rc;
    const int s = (int)(p->host_steps & 1);
    const bool reuse = p->host_steps >= 2;
    cudaStream_t cs = p->host_stream;
    // copy-in: the slot's H is free once the aggregation two steps ago has read it
    if (reuse) CU(p, cudaStreamWaitEvent(p->s_in, p->ev_comp[s], 0));
    CU(p, cudaMemcpyAsync(p->d_hostH[s], H_host, (size_t)need * 4, cudaMemcpyHostToDevice, p->s_in));
    CU(p, cudaEventRecord(p->ev_in[s], p->s_in));
    // compute: needs this step's H, and the slot's Z must have left for the host (two steps ago)
    CU(p, cudaStreamWaitEvent(cs, p->ev_in[s], 0));
    if (reuse) CU(p, cudaStreamWaitEvent(cs, p->ev_out[s], 0));
    if ((rc = pgcn_forward(p, p->d_hostH[s], p->d_hostZ[s], f, cs))) return rc;
    CU(p, cudaEventRecord(p->ev_comp[s], cs));
    // copy-out
    CU(p, cudaStreamWaitEvent(p->s_out, p->ev_comp[s], 0));
    CU(p, cudaMemcpyAsync(Z_host, p->d_hostZ[s], (size_t)need * 4, cudaMemcpyDeviceToHost, p->s_out));
    CU(p, cudaEventRecord(p->ev_out[s], p->s_out));
    ++p->host_steps;
    return 0;
}

int pgcn_forward_host_wait(pgcn_plan* p)
{
    if (!p) return fail(nullptr, PGCN_ERR_INVALID, "null plan");
    if (!p->s_out) return 0;
    CU(p, cudaStreamSynchronize(p->s_out));
    CU(p, cudaStreamSynchronize(p->host_stream));
    return 0;
}

int pgcn_forward_host(pgcn_plan* p, const float* H_host, float* Z_host, int32_t f)
{
    int rc = pgcn_forward_host_async(p, H_host, Z_host, f);
    if (rc) return rc;
    return pgcn_forward_host_wait(p);
}

}  // extern "C"
