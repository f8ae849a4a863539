// spmm_ring.cuh — the Blackwell-native SpMM of the PGCN aggregation path: gathered H rows are staged
// ASYNCHRONOUSLY in shared memory (1-D TMA bulk copies, cp.async.bulk + mbarrier complete_tx), the
// segmented FMA runs out of shared memory.
//
// Replaces torch.sparse.mm(A, H) / torch.sparse.mm(A.t(), g) of GPU/PGCN.py:127,132 for feature widths
// that are multiples of 128 floats (the benchmark widths 128 and 256); other widths take the register
// pipeline of spmm_kernels.cuh.
//
// Why: the register-buffered gather of spmm_rowblock_kernel keeps bytes-in-flight in REGISTERS (2 rows per
// warp x 48 warps), spends ~35 issue slots per edge and is capped by occupancy. Here
//   * every WARP owns a private ring of NS row slots (NS x f x 4 bytes of shared memory) and a small ring
//     of index pieces; nothing is shared between warps, so there is no CTA-level synchronisation at all;
//   * lanes 0..G-1 each issue ONE bulk copy (UBLKCP) of a whole H row (512 B at f = 128) per group of G
//     edges; the group's mbarrier completes when all G rows have landed (expect_tx = G x row bytes);
//   * the (column|flags, value) stream is fetched the same way in 256-byte pieces of 32 entries, so the
//     kernel issues no ordinary global loads at all on its hot path;
//   * consumption is one broadcast LDS.64 (index pair) + one conflict-free LDS.128 per edge and lane,
//     4 FFMA, a row-end test; bytes in flight per SM = warps x NS x row bytes (e.g. 12 x 32 x 512 B =
//     192 KB) instead of 48 KB, at ~10 issue slots per edge.
//   * MODE 1 is the same ring filled with per-lane 16-byte cp.async (LDGSTS) copies — kept for comparison.
// Row blocks come from the same host schedule as the register kernel (rows cut into blocks of about
// edges_per_block entries, long rows split into single-row segments that the fixup kernel sums in a fixed
// order), one block per warp; with `counter` set, the CTAs are persistent and warps fetch blocks dynamically.
#pragma once
#include <cuda.h>
#include "spmm_kernels.cuh"

namespace pgcn {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

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
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier; L2 eviction policy per copy.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar,
                                         unsigned long long pol)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
// 2-D tensor-map TMA in tile::gather4 mode: FOUR arbitrary rows of H (row indices r0..r3, column offset x) land as
// four consecutive row slots in shared memory with ONE instruction (UTMALDG.2D.GATHER4).
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* tm, int x, int r0, int r1, int r2, int r3,
                                            uint32_t bar, unsigned long long pol)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3, %4, %5, %6}], [%7], %8;"
        ::"r"(dst), "l"(tm), "r"(x), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, unsigned long long pol)
{
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One call site per row end would replicate the store sequence 32+ times in the unrolled consumer; keeping it
// out of line keeps the hot loop inside the instruction cache (the rows-end path runs once per ~17 edges).
template <int VPL>
__device__ __noinline__ void ring_store_row(float4 a0, float4 a1, const int* rowids, int row, float* Z0, float* Z1,
                                            int zsplit, size_t pitch, size_t off, int beta, int relu, const unsigned char* final)
{
    const bool relu_row = relu && (final == nullptr || __ldg(final + row) != 0);
    const int orow = (rowids != nullptr) ? __ldg(rowids + row) : row;
    char* zb = (orow < zsplit) ? reinterpret_cast<char*>(Z0) + (size_t)(unsigned)orow * pitch
                               : reinterpret_cast<char*>(Z1) + (size_t)(unsigned)(orow - zsplit) * pitch;
    zb += off;
    float4* zp = reinterpret_cast<float4*>(zb);
    if (beta) vadd(a0, *zp);
    if (relu_row) a0 = vrelu(a0);
    st_out(zp, a0);
    if (VPL == 2) {
        float4* zq = reinterpret_cast<float4*>(zb + 512);
        if (beta) vadd(a1, *zq);
        if (relu_row) a1 = vrelu(a1);
        st_out(zq, a1);
    }
}

constexpr int kRingWarps = 2;        // warps per CTA (each fully independent: CTA size only sets the smem granule)
constexpr int kRingPieces = 4;       // index pieces (32 entries, kPieceBytes each) resident per warp

struct RingArgs {
    unsigned int* counter;   // null: block = blockIdx.x * kRingWarps + warp; else dynamic (persistent CTAs)
    const float* hub;        // reserved (hub rows resident in shared memory)
    int nhub;
};

// per warp: NS row slots | NP index pieces | NG + NP mbarriers
__host__ __device__ constexpr size_t ring_warp_bytes(int vpl, int ns, int ng)
{
    return ((size_t)ns * vpl * 512 + (size_t)kRingPieces * kPieceBytes + (size_t)(ng + kRingPieces) * 8 + 127) / 128 * 128;
}
__host__ __device__ constexpr size_t ring_smem_bytes(int vpl, int ns, int ng)
{
    return ring_warp_bytes(vpl, ns, ng) * kRingWarps + 128;
}

// runtime-indexed access to a tiny register array (compare chain instead of local memory)
template <int N>
__device__ __forceinline__ uint32_t reg_get(const uint32_t (&a)[N], int i)
{
    uint32_t v = a[0];
#pragma unroll
    for (int k = 1; k < N; ++k) v = (i == k) ? a[k] : v;
    return v;
}
template <int N>
__device__ __forceinline__ void reg_set(uint32_t (&a)[N], int i, uint32_t v)
{
#pragma unroll
    for (int k = 0; k < N; ++k) a[k] = (i == k) ? v : a[k];
}

// VPL : 128-float vector groups per row (tile of f).
// G   : edges per completion group (8, 16 or 32); NG: groups in the ring (2 or 4); the warp owns NS = G * NG row slots.
// MODE: 0 = 1-D TMA bulk copies (UBLKCP, one per row), 1 = per-lane 16-byte cp.async (LDGSTS + wait_group),
//       2 = 2-D tensor-map TMA in tile::gather4 mode (UTMALDG.2D.GATHER4, FOUR rows per instruction; quads that mix
//           own and halo columns, or groups cut by a block boundary, fall back to the 1-D copies of MODE 0).
// HALO: columns >= split live in a second matrix (the halo slab).
//
// The entries are walked in GLOBALLY ALIGNED units: a piece = entries [32 P, 32 P + 32) (one 272-byte record =
// one bulk copy: 32 plain column indices, 32 values, a row-end bit mask and a cold-column bit mask), a group =
// entries [G g, G g + G) (one completion unit of the row ring, slot group g % NG). A row block [e0, e1) starts and
// ends anywhere; entries of its first / last group outside the block are masked. The steady state per group is:
//   issue   : one broadcast LDS.64 (masks), lanes 0..G/4-1 read their four column indices with one LDS.128 and
//             fire one gather4 each; lane 0 arms the group's mbarrier with G x row bytes;
//   consume : mbarrier wait, 8 x LDS.128 rows + 2 x LDS.128 values (straight from the resident piece), 32 FFMA per
//             8 edges; a row-end test per edge only in groups whose row-end mask is non-zero.
// Pieces stay resident until their last group has been CONSUMED (the values are read at consumption time), so
// nothing is copied between issue and consumption except the row-end mask, which travels in a register.
template <int VPL, int G, int NG, int MODE, bool HALO>
__device__ __forceinline__ void ring_body(const SpmmArgs& a, const RingArgs& ra, const CUtensorMap* tm0, const CUtensorMap* tm1)
{
    constexpr int NS = G * NG, NP = kRingPieces;
    constexpr int PG = 32 / G;                                           // groups per index piece
    constexpr int U = (NG > PG) ? NG / PG : 1;                           // pieces per loop body (body groups % NG == 0)
    constexpr uint32_t RB = VPL * 512;                                   // bytes of one row tile
    constexpr uint32_t FULL = (G == 32) ? 0xffffffffu : ((1u << G) - 1u);
    static_assert((G == 8 || G == 16 || G == 32) && (NG == 2 || NG == 4), "unsupported ring shape");
    static_assert((U * PG) % NG == 0, "slot groups must repeat with the loop body");
    static_assert(NP * 32 >= NS + 64, "pieces must stay resident from prefetch to consumption");
    extern __shared__ __align__(128) unsigned char ring_smem[];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wbase = ring_smem + (size_t)warp * ring_warp_bytes(VPL, NS, NG);
    const uint32_t s_data = smem_u32(wbase);                             // NS slots of RB bytes
    const uint32_t s_idx = s_data + NS * RB;                             // NP pieces
    const uint32_t s_gbar = s_idx + NP * kPieceBytes;                    // NG group barriers
    const uint32_t s_pbar = s_gbar + NG * 8;                             // NP piece barriers
    const int* idx_gen = reinterpret_cast<const int*>(wbase + (size_t)NS * RB);
    const float4* data_gen = reinterpret_cast<const float4*>(wbase) + lane;

    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NG + NP; ++i) mbar_init(s_gbar + i * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();

    const unsigned long long pol_hot = l2_policy_evict_last();
    const unsigned long long pol_cold = l2_policy_evict_first();
    // blockIdx.y walks feature tiles of 128 * VPL floats (f = 384, 512, ...): row pitch f * 4, tile offset y * RB
    const size_t pitch = (size_t)a.f * 4;
    const size_t toff = (size_t)blockIdx.y * RB;
    const char* hb0 = reinterpret_cast<const char*>(a.H0) + toff;
    const char* hb1 = HALO ? reinterpret_cast<const char*>(a.H1) + toff - (size_t)a.split * pitch : hb0;
    const unsigned usplit = HALO ? (unsigned)a.split : 0xffffffffu;
    unsigned int* counter = ra.counter ? ra.counter + blockIdx.y : nullptr;

    uint32_t gpar = 0;                   // phase parity of each group barrier (bit sg)
    // pieces move through the NP slots as a FIFO: fetched -> landed (issue side) -> consumed
    uint32_t pfetch = 0, pwait = 0, pcons = 0;

    float4 acc[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);

    int blk = counter ? 0 : (int)(blockIdx.x * kRingWarps + warp);
    if (counter) {
        if (lane == 0) blk = (int)atomicAdd(counter, 1u);
        blk = __shfl_sync(0xffffffffu, blk, 0);
    }

    while (blk < a.nblocks) {
        const int4 b = __ldg(a.blocks + blk);
        const bool seg = b.y < 0;
        const int e0 = b.z, e1 = b.w;
        int row = b.x;
        const int gA = e0 / G, gB = (e1 - 1) / G;                        // first / last (aligned) group
        const int P0 = e0 >> 5, P1 = (e1 - 1) >> 5;                      // first / last piece
        uint32_t vmask[NG], emask[NG];                                   // per slot group: valid edges, row ends
#pragma unroll
        for (int i = 0; i < NG; ++i) vmask[i] = emask[i] = 0;
        int pnext = P0;                                                  // next piece to fetch
        const int* pi = idx_gen;                                         // piece being issued from
        const int* pc = idx_gen + (pcons % NP) * kPieceInts;             // piece being consumed from

        auto fetch_piece = [&]() {
            if (pnext <= P1) {
                if (lane == 0) {
                    const uint32_t q = pfetch % NP;
                    mbar_expect_tx(s_pbar + q * 8, kPieceBytes);
                    bulk_g2s(s_idx + q * kPieceBytes, a.pieces + (size_t)pnext * kPieceInts, kPieceBytes, s_pbar + q * 8, pol_cold);
                }
                ++pfetch;
                ++pnext;
            }
        };
        auto wait_piece = [&]() {                                        // the next piece in FIFO order has landed
            const uint32_t q = pwait % NP;
            mbar_wait(s_pbar + q * 8, (pwait / NP) & 1);
            pi = idx_gen + q * kPieceInts;
            ++pwait;
        };
        auto piece_consumed = [&]() {                                    // the consumer leaves its piece: slot is free
            ++pcons;
            pc = idx_gen + (pcons % NP) * kPieceInts;
            fetch_piece();
        };
        auto flush_row = [&]() {
            ring_store_row<VPL>(acc[0], acc[VPL - 1], a.rowids, row, a.Z0, a.Z1, a.zsplit, pitch, toff + lane * 16, a.beta, a.relu, a.final);
#pragma unroll
            for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
            ++row;
        };
        // issue the group at sub-position qs of piece `pi` into slot group sg; vm = valid entries (FULL inside the block)
        auto issue = [&](int qs, int sg, uint32_t vm) {
            const uint2 m = *reinterpret_cast<const uint2*>(pi + 64);    // {row-end mask, cold mask} of the piece
            const uint32_t em = seg ? 0u : ((m.x >> (qs * G)) & vm);
            const uint32_t cm = (m.y >> (qs * G)) & FULL;
            reg_set(emask, sg, em);
            reg_set(vmask, sg, vm);
            const int* cols = pi + qs * G;
            if (MODE == 1) {
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    if (vm >> j & 1) {
                        const unsigned cj = (unsigned)cols[j];
                        const char* src = (cj >= usplit ? hb1 : hb0) + (size_t)cj * pitch + lane * 16;
#pragma unroll
                        for (int v = 0; v < VPL; ++v)
                            cp_async16(s_data + (sg * G + j) * RB + v * 512 + lane * 16, src + v * 512,
                                       (cm >> j & 1) ? pol_cold : pol_hot);
                    }
                }
                cp_async_commit();
                return;
            }
            if (lane == 0) mbar_expect_tx(s_gbar + sg * 8, (uint32_t)__popc(vm) * RB);
            uint32_t single = vm;                                        // entries copied one row at a time (1-D bulk)
            if (MODE == 2 && vm == FULL) {
                bool fire = lane < G / 4;
                int4 q4 = make_int4(0, 0, 0, 0);
                if (fire) q4 = *reinterpret_cast<const int4*>(cols + lane * 4);
                bool allhalo = false;
                if (HALO) {
                    const unsigned r0 = (unsigned)q4.x, r1 = (unsigned)q4.y, r2 = (unsigned)q4.z, r3 = (unsigned)q4.w;
                    const bool allown = r0 < usplit && r1 < usplit && r2 < usplit && r3 < usplit;
                    allhalo = r0 >= usplit && r1 >= usplit && r2 >= usplit && r3 >= usplit;
                    fire = fire && (allown || allhalo);
                    const uint32_t okq = __ballot_sync(0xffffffffu, fire);      // bit q: quad q goes out as one gather4
                    single = 0;
#pragma unroll
                    for (int q = 0; q < G / 4; ++q) single |= (okq >> q & 1) ? 0u : (0xfu << (4 * q));
                } else {
                    single = 0;
                }
                if (fire) {
                    const int sub = allhalo ? (int)usplit : 0;
                    const bool cold = ((cm >> (lane * 4)) & 0xfu) == 0xfu;
                    tma_gather4(s_data + (sg * G + lane * 4) * RB, allhalo ? tm1 : tm0, (int)(blockIdx.y * (RB / 4)),
                                q4.x - sub, q4.y - sub, q4.z - sub, q4.w - sub, s_gbar + sg * 8, cold ? pol_cold : pol_hot);
                }
            }
            if (lane < G && (single >> lane & 1)) {
                const unsigned cj = (unsigned)cols[lane];
                bulk_g2s(s_data + (sg * G + lane) * RB, (cj >= usplit ? hb1 : hb0) + (size_t)cj * pitch, RB,
                         s_gbar + sg * 8, (cm >> lane & 1) ? pol_cold : pol_hot);
            }
        };
        // consume slot group sg = the group at sub-position qs of piece `pc`
        auto consume = [&](int qs, int sg, bool interior) {
            const uint32_t vm = interior ? FULL : reg_get(vmask, sg);
            if (vm == 0) {                                               // group outside the block: nothing was issued
                if (MODE == 1) cp_async_wait<NG - 1>();
                return;
            }
            if (MODE != 1) { mbar_wait(s_gbar + sg * 8, (gpar >> sg) & 1); gpar ^= 1u << sg; }
            else cp_async_wait<NG - 1>();
            const uint32_t em = reg_get(emask, sg);
            const float4* slot = data_gen + (size_t)(sg * G) * (RB / 16);
            const float* wv = reinterpret_cast<const float*>(pc) + 32 + qs * G;
            if (vm == FULL) {
#pragma unroll
                for (int c = 0; c < G; c += 8) {                         // 8 rows at a time: 8 x LDS.128 in flight
                    const float4 wa = *reinterpret_cast<const float4*>(wv + c);
                    const float4 wb = *reinterpret_cast<const float4*>(wv + c + 4);
                    const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
                    float4 r[8][VPL];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
#pragma unroll
                        for (int v = 0; v < VPL; ++v) r[j][v] = slot[(c + j) * (RB / 16) + v * 32];
                    if (((em >> c) & 0xffu) == 0) {                      // no row end inside: 8 x (LDS.128, 4 FFMA)
#pragma unroll
                        for (int j = 0; j < 8; ++j)
#pragma unroll
                            for (int v = 0; v < VPL; ++v) vfma(acc[v], w[j], r[j][v]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
#pragma unroll
                            for (int v = 0; v < VPL; ++v) vfma(acc[v], w[j], r[j][v]);
                            if (em >> (c + j) & 1) flush_row();
                        }
                    }
                }
            } else {                                                     // first / last group of a block
#pragma unroll 1
                for (int j = 0; j < G; ++j) {
                    if (vm >> j & 1) {
                        const float wj = wv[j];
#pragma unroll
                        for (int v = 0; v < VPL; ++v) vfma(acc[v], wj, slot[j * (RB / 16) + v * 32]);
                        if (em >> j & 1) flush_row();
                    }
                }
            }
            __syncwarp();                                                // every lane is done with these slots
        };
        // valid-entry mask of aligned group gi with respect to the block [e0, e1)
        auto group_mask = [&](int gi) -> uint32_t {
            if (gi < gA || gi > gB) return 0u;
            uint32_t vm = FULL;
            if (gi == gA) vm &= FULL << (e0 - gA * G);
            if (gi == gB) vm &= FULL >> (G - 1 - ((e1 - 1) - gB * G));
            return vm;
        };
        auto issue_checked = [&](int gi, int qs, int sg) {
            const uint32_t vm = group_mask(gi);
            if (vm == 0) { reg_set(vmask, sg, 0u); if (MODE == 1) cp_async_commit(); return; }
            issue(qs, sg, vm);
        };

        // prologue: index pieces in flight, first piece landed, the first NG groups issued
#pragma unroll
        for (int i = 0; i < NP; ++i) fetch_piece();
        wait_piece();
#pragma unroll 1
        for (int i = 0; i < NG; ++i) {
            if (i > 0 && i % PG == 0 && P0 + i / PG <= P1) wait_piece();     // the ring spans more than one piece
            if (P0 + i / PG <= P1) issue_checked(PG * P0 + i, i % PG, i);
            else { reg_set(vmask, i, 0u); if (MODE == 1) cp_async_commit(); }
        }

        // The loops below are deliberately NOT unrolled over the groups of a piece: slot group and sub-group are
        // runtime values (a handful of integer instructions per group), which keeps the whole kernel inside the
        // instruction cache; the 8-row consume chunks inside a group are unrolled.
        for (int P = P0; P <= P1; P += U) {
            // interior: every group consumed AND every group issued by this body lies strictly inside the block
            const bool interior = PG * P > gA && PG * (P + U) - 1 + NG < gB;
#pragma unroll 1
            for (int idx = 0; idx < U * PG; ++idx) {
                const int sg = idx % NG;
                const int qc = idx % PG;                                 // sub-position of the consumed group
                consume(qc, sg, interior);
                if (qc == PG - 1) piece_consumed();
                const int qi = (idx + NG) % PG;                          // sub-position of the group NG ahead
                const int Pi = P + (idx + NG) / PG;                      // its piece
                if (interior) {
                    if (qi == 0) wait_piece();
                    issue(qi, sg, FULL);
                } else {
                    if (qi == 0 && Pi <= P1) wait_piece();               // first group of a new piece
                    if (Pi <= P1) issue_checked(PG * Pi + qi, qi, sg);
                    else { reg_set(vmask, sg, 0u); if (MODE == 1) cp_async_commit(); }
                }
            }
        }
        // pieces P1+1 .. (rounded up to the body) were never fetched, but the body counted them as consumed
        pcons = pwait;

        if (seg) {
            char* pb = reinterpret_cast<char*>(a.partial) + (size_t)(unsigned)(-b.y - 1) * pitch + toff;
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                reinterpret_cast<float4*>(pb + v * 512)[lane] = acc[v];
                acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }

        if (!counter) break;
        if (lane == 0) blk = (int)atomicAdd(counter, 1u);
        blk = __shfl_sync(0xffffffffu, blk, 0);
    }
    if (MODE == 1) cp_async_wait<0>();
}

template <int VPL, int G, int NG, int MODE, bool HALO>
__global__ void __launch_bounds__(kRingWarps * 32)
spmm_ring_kernel(const SpmmArgs a, const RingArgs ra)
{
    ring_body<VPL, G, NG, MODE, HALO>(a, ra, nullptr, nullptr);
}

// gather4 variant: the tensor maps of H_own (tm0) and of the halo slab (tm1) travel as __grid_constant__ parameters
template <int VPL, int G, int NG, bool HALO>
__global__ void __launch_bounds__(kRingWarps * 32)
spmm_ring_g4_kernel(const SpmmArgs a, const RingArgs ra, const __grid_constant__ CUtensorMap tm0,
                    const __grid_constant__ CUtensorMap tm1)
{
    ring_body<VPL, G, NG, 2, HALO>(a, ra, &tm0, &tm1);
}

}  // namespace pgcn
