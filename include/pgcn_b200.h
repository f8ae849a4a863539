/*
 * pgcn_b200.h — C-ABI of the B200-native PGCN aggregation path.
 *
 * This is the drop-in boundary for ONE hot path of the reference
 * (gunduzvd/Scalable-Graph-Convolutional-Network-Training-on-Distributed-Memory-Systems):
 * the per-layer sparse neighbour aggregation  Z = A_local * H  plus the boundary-row
 * (halo) exchange that the 1-D row partition induces.  In the reference that path is
 *
 *     GPU/PGCN.py:85-119   communicate_fgm(H, backward)     (pack / send / recv / unpack / H+X)
 *     GPU/PGCN.py:121-134  PSpMM.forward / PSpMM.backward   (torch.sparse.mm(A, H), torch.sparse.mm(A.t(), g))
 *     GPU/PGCN.py:37-64    compute_communication_maps, get_partitiont_of_adjacency_matrix (plan inputs)
 *     Parallel-GCN/main.c:269-299 / :374-404   GrB_mxm PLUS_TIMES_FP32 aggregation (CPU twin)
 *
 * Conventions
 *   - every entry point is extern "C", returns 0 on success or a negative pgcn_status,
 *     never throws, never calls exit(); the text of the last error is pgcn_last_error().
 *   - index arrays handed to pgcn_plan_create are HOST pointers and are copied; feature
 *     matrices (H, Z, G, slabs) are DEVICE pointers owned by the caller (torch), row-major
 *     fp32 with a leading dimension equal to f.
 *   - all compute entry points are asynchronous on the cudaStream_t passed in (as void*).
 *   - no torch types appear here.
 *
 * Column space of the local matrix: rank r renumbers columns to
 *     [ own rows (m) | halo rows received from peer 0 | ... | from peer k-1 ]
 * each halo group sorted by global vertex id — the same order as the reference's
 * sorted send/recv maps (GPU/PGCN.py:47-48), so sender order == receiver order.
 */
#ifndef PGCN_B200_H
#define PGCN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pgcn_plan pgcn_plan;

typedef enum pgcn_status {
    PGCN_OK = 0,
    PGCN_ERR_INVALID = -1,   /* bad argument (null pointer, negative size, f > f_max, ...) */
    PGCN_ERR_CUDA = -2,      /* a CUDA runtime call failed                                   */
    PGCN_ERR_NCCL = -3,      /* NCCL missing or an NCCL call failed                          */
    PGCN_ERR_NOGPU = -4,     /* no CUDA device visible: there is NO CPU fallback             */
    PGCN_ERR_STATE = -5      /* call made in the wrong state (no communicator, ...)          */
} pgcn_status;

/* Byte counts for the roofline (SURVEY.md §8d "Algorithmic bytes"). All per call, this rank. */
typedef struct pgcn_bytes {
    int64_t nnz;            /* nnz_loc of the forward matrix                                  */
    int64_t m;              /* owned rows                                                      */
    int64_t h;              /* halo rows                                                       */
    int64_t cols_ref;       /* distinct columns referenced (own + halo)                        */
    int64_t spmm_fwd;       /* 8*nnz + 4*(m+1) + 4*f*cols_ref + 4*f*m                          */
    int64_t spmm_bwd;       /* 8*nnz + 4*(m+h+1) + 4*f*rows_ref_t + 4*f*(m+h)                  */
    int64_t gather_fwd;     /* no-reuse bound: nnz*(8+4f) + 4*(m+1) + 4*f*m                    */
    int64_t xchg_out;       /* 4*f*S   bytes this rank sends in a forward exchange             */
    int64_t xchg_in;        /* 4*f*h   bytes this rank receives in a forward exchange          */
    int64_t pack;           /* 2*4*f*S HBM bytes of the pack kernel                            */
} pgcn_bytes;

/* ---- library-level ---------------------------------------------------------------------- */

/* Version / build string, e.g. "pgcn_b200 0.1 sm_100a". Never NULL. */
const char* pgcn_version(void);

/* Number of visible CUDA devices, or a negative pgcn_status. */
int pgcn_device_count(void);

/* Text of the last error raised on `plan` (or on the library when plan == NULL). Never NULL. */
const char* pgcn_last_error(const pgcn_plan* plan);

/* ---- plan: replaces GPU/PGCN.py:37-64 (maps + local matrix) and :178-182 (buffers) ------- */

/*
 * Build the per-rank plan on the CURRENT CUDA device.
 *   rowptr/colidx/vals        forward CSR of the owned rows: m rows, columns in [0, m+h)
 *   t_rowptr/t_colidx/t_vals  CSR of its transpose: m+h rows, columns in [0, m)
 *                             (replaces the per-call A.t() + re-coalesce of GPU/PGCN.py:132)
 *   send_idx                  S = send_off[k] local row ids; rows for peer p are
 *                             send_idx[send_off[p] .. send_off[p+1])   (GPU/PGCN.py:47 send_map)
 *   recv_off                  halo rows from peer p live at columns m+recv_off[p] .. m+recv_off[p+1);
 *                             h = recv_off[k]                          (GPU/PGCN.py:48 recv_map)
 *   f_max                     largest feature width that will be used (sizes the slabs)
 * Duplicated (row, col) entries are allowed and are summed, like the uncoalesced COO of
 * GPU/PGCN.py:60-63.
 */
int pgcn_plan_create(const int32_t* rowptr, const int32_t* colidx, const float* vals,
                     int32_t m, int32_t h,
                     const int32_t* t_rowptr, const int32_t* t_colidx, const float* t_vals,
                     const int32_t* send_idx, const int64_t* send_off, const int64_t* recv_off,
                     int32_t k, int32_t rank, int32_t f_max,
                     pgcn_plan** out);

int pgcn_plan_destroy(pgcn_plan* plan);

/*
 * Plan options (take effect at the next compute call). Names:
 *   "kernel"               0 = automatic (default): widths that are multiples of 128 floats with 16-byte aligned
 *                          operands take the shared-memory ring kernel fed by TMA tile::gather4, everything else the
 *                          register-pipeline kernel; 4 = always the register kernel; 5 / 6 / 7 = ring kernel fed by
 *                          1-D cp.async.bulk / cp.async / tile::gather4
 *   "ring_slots"           row slots per warp of the ring kernel: 16 (default), 32, 64; "ring_groups" 2 | 4 (64 slots)
 *   "ring_edges_per_block" target nnz of one row block = one warp's unit of work          (default 512)
 *   "ring_long_row"        rows with more nnz than this are split into segments           (default 2 * block)
 *   "persistent"           1 = persistent CTAs fetch row blocks dynamically (default; single-rank plans and plans
 *                          without overlap), "persistent_multi" 1 = also for overlapped multi-rank plans (default 0:
 *                          persistent CTAs would hold the SMs the exchange kernels need)
 *   "edges_per_block", "long_row", "tile_floats"   the same for the register kernel (defaults 128, 4 * block, 0)
 *   "overlap"              1 = split A_local into own / per-peer halo blocks and pipeline the exchange with them
 *                          (Parallel-GCN/main.c:271 then :275-299)                         (default 1)
 *   "relu"                 1 = pgcn_forward writes max(0, A_local * H)                      (default 0)
 *   "p2p"                  0 = never use the peer-memory transport (all ranks must agree)  (default 1)
 * pgcn_plan_autotune overrides block sizes / ring depth per matrix; setting one of them explicitly clears the tuned
 * values. "hot_mb" (L2-resident hot set of H rows) is fixed at plan creation: environment variable PGCN_HOT_MB.
 * Split rows are always reduced in a fixed order: results are run-to-run deterministic.
 * Read-only names for pgcn_plan_get_option: "nccl", "blocks_fwd", "long_rows_fwd", "ring_blocks_fwd",
 * "ring_long_rows_fwd".
 */
int pgcn_plan_set_option(pgcn_plan* plan, const char* name, int64_t value);
int64_t pgcn_plan_get_option(const pgcn_plan* plan, const char* name);

/*
 * Time the forward SpMM of this plan at feature width f for a few "edges_per_block" values on
 * scratch buffers and keep the fastest (set-up work, like the reference's untimed plan building,
 * GPU/PGCN.py:171-200). Synchronous. Returns the chosen value (>0) or a negative pgcn_status.
 */
int pgcn_plan_autotune(pgcn_plan* plan, int32_t f);

/*
 * Host-only (no GPU needed): the row-block schedule the SpMM walks, for inspection and tests.
 * rowptr must describe NON-EMPTY rows only (the plan squeezes empty rows out first). Writes up to cap_blocks
 * blocks as 4 int32 each {first row, nrows | -(slot+1), e_begin, e_end}; returns the number of blocks.
 */
int64_t pgcn_debug_schedule(const int32_t* rowptr, int32_t nrows, int64_t edges_per_block, int64_t long_row,
                            int32_t* blocks_out, int64_t cap_blocks, int32_t* nlong_out, int32_t* nslots_out);

/* Plan-owned device slabs (f_max floats per row), for callers that want zero-copy access:
 * which = 0 send slab (S rows), 1 halo/recv slab (h rows), 2 reverse recv slab (S rows),
 * 3 reverse send slab (h rows: halo partials of A^T g). */
void* pgcn_plan_slab(pgcn_plan* plan, int which);

int pgcn_algorithmic_bytes(const pgcn_plan* plan, int32_t f, pgcn_bytes* out);

/* Number of kernels launched by this plan since creation (bench.py's "gpu_launches"). */
int64_t pgcn_launch_count(const pgcn_plan* plan);

/* ---- communicator: replaces dist.init_process_group + dist.send/recv (GPU/PGCN.py:107,112,242) */

/* Fill `id128` (128 bytes) with an NCCL unique id (rank 0 calls this, then ships the bytes). */
int pgcn_comm_unique_id(void* id128);
/* Collective over the k ranks of the plan. */
int pgcn_comm_init(pgcn_plan* plan, const void* id128);

/*
 * Peer-memory transport (single NVSwitch box): the pack kernel stores boundary rows straight
 * into the peer's halo slab over NVLink, no staging copy and no NCCL on the data path.
 *   pgcn_p2p_export : writes PGCN_P2P_HANDLE_BYTES: this rank's CUDA IPC handle for its exchange arena + layout
 *   pgcn_p2p_import : takes the k handles/layouts gathered from all ranks (rank-major)
 */
#define PGCN_P2P_HANDLE_BYTES 512
/* Many plans, one communicator (mini-batch training: one plan per pre-sampled batch, GPU/PGCN-Mini-batch.py:220-230
 * swaps [bA, send_map, recv_map] per batch over the same process group): `plan` borrows `owner`'s NCCL communicator
 * (same rank / size / device); the owner must outlive the borrowers. */
int pgcn_comm_share(pgcn_plan* plan, pgcn_plan* owner);
int pgcn_p2p_export(pgcn_plan* plan, void* handle_out);
int pgcn_p2p_import(pgcn_plan* plan, const void* handles_k);

/* ---- the hot path ------------------------------------------------------------------------ */

/*
 * Z = op(A_local) * [H_own ; H_halo]         (GPU/PGCN.py:127 and :132 without the exchange)
 *   transpose = 0 : A (m rows). H_own is m x f, H_halo is h x f (may be NULL when h == 0).
 *                   Z is m x f; Z_halo is ignored.
 *   transpose = 1 : A^T (m+h rows). H_own is the m x f upstream gradient, H_halo ignored.
 *                   Rows [0,m) go to Z (m x f), rows [m, m+h) go to Z_halo (h x f), already in
 *                   the wire order of the reverse exchange.
 *   transpose = 2 : own-columns half of the overlapped forward, Z  = A_own  * H_own
 *   transpose = 3 : halo-columns half,                          Z += A_halo * H_halo
 *                   (Parallel-GCN/main.c:271 then :295; plans with k > 1 and h > 0 only)
 */
int pgcn_spmm(pgcn_plan* plan, int transpose,
              const float* H_own, const float* H_halo,
              float* Z, float* Z_halo, int32_t f, void* stream);

/* send_slab[j, :] = H[send_idx[j], :]  for all peers in one launch   (GPU/PGCN.py:104) */
int pgcn_pack(pgcn_plan* plan, const float* H, float* send_slab, int32_t f, void* stream);

/*
 * All-to-all-v of row slabs (GPU/PGCN.py:99-115, both phases, all peers, one grouped call).
 *   reverse = 0 : send send_slab rows [send_off[p], send_off[p+1]) to p; receive halo rows
 *                 [recv_off[p], recv_off[p+1]) from p.
 *   reverse = 1 : the gradient direction (maps swapped, GPU/PGCN.py:93-97).
 * Zero-length messages are skipped on the wire (they are still counted by the host stats).
 */
int pgcn_exchange(pgcn_plan* plan, const float* send_slab, float* recv_slab,
                  int32_t f, int reverse, void* stream);

/*
 * G_own[send_idx[j], :] += recv_slab[j, :]   summed over ALL j in a fixed order
 * (the intended semantics of GPU/PGCN.py:115 in backward; the reference ASSIGNS — quirk Q3).
 */
int pgcn_unpack_add(pgcn_plan* plan, const float* recv_slab, float* G_own, int32_t f, void* stream);

/*
 * Fused convenience entry points = PSpMM.forward / PSpMM.backward (GPU/PGCN.py:123-134).
 *   forward : pack -> exchange (NCCL, or peer stores when p2p was imported) overlapped with
 *             the own-columns part of the SpMM -> halo-columns part. Z (m x f) = A_local * H.
 *   backward: G (m x f) = (A_local^T * gZ)[own] + contributions received from the peers.
 * With k == 1 no communicator is needed.
 */
/* Plan option "relu" = 1 fuses the layer epilogue into the forward: Z = max(0, A_local * H), clamped in the store of
 * whichever launch writes a row last (no extra pass over Z). With the dense step applied first — relu(A (H W)) — it is
 * the reference layer relu(linear(PSpMM(A, H))) of GPU/PGCN.py:144-148 up to fp32 association. Forward only. */
int pgcn_forward(pgcn_plan* plan, const float* H_own, float* Z, int32_t f, void* stream);
int pgcn_backward(pgcn_plan* plan, const float* gZ, float* G_own, int32_t f, void* stream);

/* ---- host-buffer variant: what a non-torch host (the reference's C path) would bind -------- */
/*
 * Same as pgcn_forward but H and Z are HOST pointers (pinned or pageable): copies H to the
 * device, runs the path, copies Z back, synchronises. This is the call bench.py times as "e2e".
 */
int pgcn_forward_host(pgcn_plan* plan, const float* H_host, float* Z_host, int32_t f);
/*
 * Software-pipelined form for a host that streams many aggregations (layers, mini-batches, timesteps):
 * pgcn_forward_host_async enqueues  H_host -> device slot (copy-in stream) ; pgcn_forward (compute stream) ;
 * device slot -> Z_host (copy-out stream)  on one of TWO device slots and returns at once, so the upload of
 * step i+1 and the download of step i-1 run under the aggregation of step i (PCIe is full duplex).
 * H_host must stay untouched and Z_host unread until pgcn_forward_host_wait returns; it waits for everything
 * enqueued so far. Pinned host buffers are needed for the copies to overlap. Replaces nothing in the reference
 * (it has no host-resident path, GPU/PGCN.py:186-196 keeps H on the device): it is the C-trainer-facing form of
 * the PSpMM.forward boundary (GPU/PGCN.py:123-127).
 */
int pgcn_forward_host_async(pgcn_plan* plan, const float* H_host, float* Z_host, int32_t f);
int pgcn_forward_host_wait(pgcn_plan* plan);

#ifdef __cplusplus
}
#endif
#endif /* PGCN_B200_H */
