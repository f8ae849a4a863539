// part_tool — offline producer of the benchmark part vectors (hp / gp), test & bench infrastructure only.
//
// It drives the SAME partitioning libraries with the SAME models and parameters as the reference's
// partitioner drivers, linked against the binaries vendored in the reference tree where they lie
// (never copied):
//   hp: column-net hypergraph, PaToH connectivity-1 metric, cell weight = stored entries of the row,
//       unit net weights, final_imbal 0.1                       /root/reference/GPU/hypergraph/main.cpp:312-386
//   gp: METIS k-way, edge-cut objective, ufactor 1, unit vertex and edge weights, diagonal dropped
//                                                               /root/reference/GPU/graph/main.cpp:300-360
// What differs from the reference drivers is only the I/O: they parse MatrixMarket text into nested
// std::unordered_map (tens of GB and minutes at 1e8 entries); this tool reads a binary CSR written by
// tools/make_partvecs.py and writes one byte per vertex. `preset` selects PaToH's suggested parameter
// set: `quality` is what the reference passes (main.cpp:346); `speed` exists because QUALITY needs
// ~8 CPU-hours on the 10 M-vertex / 110 M-entry benchmark graph.
//
//   part_tool hp|gp <k> <csr.bin> <out.u8> [quality|speed|default] [seed]
//   csr.bin: int64 n, int64 nnz, int32 rowptr[n+1], int32 colidx[nnz]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#ifdef PART_TOOL_HP
#include "patoh.h"
#endif
#ifdef PART_TOOL_GP
#include "metis.h"
#endif

static bool read_all(FILE* f, void* dst, size_t bytes) { return fread(dst, 1, bytes, f) == bytes; }

int main(int argc, char** argv)
{
    if (argc < 5) {
        fprintf(stderr, "usage: %s hp|gp <k> <csr.bin> <out.u8> [quality|speed|default] [seed]\n", argv[0]);
        return 2;
    }
    const int k = atoi(argv[2]);
    const char* preset = argc > 5 ? argv[5] : "quality";
    const int seed = argc > 6 ? atoi(argv[6]) : 1;
    FILE* f = fopen(argv[3], "rb");
    if (!f) { perror(argv[3]); return 1; }
    int64_t n = 0, nnz = 0;
    if (!read_all(f, &n, 8) || !read_all(f, &nnz, 8)) { fprintf(stderr, "short header\n"); return 1; }
    std::vector<int> rowptr((size_t)n + 1), colidx((size_t)nnz);
    if (!read_all(f, rowptr.data(), (size_t)(n + 1) * 4) || !read_all(f, colidx.data(), (size_t)nnz * 4)) {
        fprintf(stderr, "short file\n");
        return 1;
    }
    fclose(f);
    std::vector<int> part((size_t)n, 0);

    if (k > 1 && !strcmp(argv[1], "hp")) {
#ifdef PART_TOOL_HP
        // cells = vertices (weight = entries of the row), net i = the entries of row i (== column i for the
        // symmetric patterns the reference preprocessing produces): main.cpp:318-340
        std::vector<int> cw((size_t)n), nw((size_t)n, 1), partw((size_t)k);
        for (int64_t i = 0; i < n; ++i) cw[(size_t)i] = rowptr[(size_t)i + 1] - rowptr[(size_t)i];
        PaToH_Parameters args;
        const int sug = !strcmp(preset, "speed") ? PATOH_SUGPARAM_SPEED
                      : !strcmp(preset, "default") ? PATOH_SUGPARAM_DEFAULT : PATOH_SUGPARAM_QUALITY;
        PaToH_Initialize_Parameters(&args, PATOH_CONPART, sug);
        args._k = k;
        args.final_imbal = 0.1;
        args.seed = seed;
        int cut = 0;
        PaToH_Alloc(&args, (int)n, (int)n, 1, cw.data(), nw.data(), rowptr.data(), colidx.data());
        PaToH_Part(&args, (int)n, (int)n, 1, 0, cw.data(), nw.data(), rowptr.data(), colidx.data(), NULL,
                   part.data(), partw.data(), &cut);
        PaToH_Free();
        printf("hp k=%d preset=%s cut(connectivity-1)=%d\n", k, preset, cut);
#else
        fprintf(stderr, "built without PaToH\n");
        return 3;
#endif
    } else if (k > 1 && !strcmp(argv[1], "gp")) {
#ifdef PART_TOOL_GP
        std::vector<idx_t> xadj((size_t)n + 1, 0), adj;
        adj.reserve((size_t)nnz);
        for (int64_t i = 0; i < n; ++i) {
            for (int e = rowptr[(size_t)i]; e < rowptr[(size_t)i + 1]; ++e)
                if (colidx[(size_t)e] != i) adj.push_back(colidx[(size_t)e]);
            xadj[(size_t)i + 1] = (idx_t)adj.size();
        }
        std::vector<idx_t> vw((size_t)n, 1), ew(adj.size(), 1), p((size_t)n, 0);
        idx_t options[METIS_NOPTIONS];
        METIS_SetDefaultOptions(options);
        options[METIS_OPTION_PTYPE] = METIS_PTYPE_KWAY;
        options[METIS_OPTION_OBJTYPE] = METIS_OBJTYPE_CUT;
        options[METIS_OPTION_UFACTOR] = 1;
        options[METIS_OPTION_SEED] = seed;
        idx_t nv = (idx_t)n, ncon = 1, np = k, edgecut = 0;
        int rc = METIS_PartGraphKway(&nv, &ncon, xadj.data(), adj.data(), vw.data(), NULL, ew.data(), &np, NULL, NULL,
                                     options, &edgecut, p.data());
        if (rc != METIS_OK) { fprintf(stderr, "METIS failed: %d\n", rc); return 1; }
        for (int64_t i = 0; i < n; ++i) part[(size_t)i] = (int)p[(size_t)i];
        printf("gp k=%d edgecut=%d\n", k, (int)edgecut);
#else
        fprintf(stderr, "built without METIS\n");
        return 3;
#endif
    } else if (k > 1) {
        fprintf(stderr, "unknown method %s\n", argv[1]);
        return 2;
    }

    // halo volume the 1-D row partition induces: sum over columns j of (#parts referencing j) - 1 when the
    // owner is among them (what GPU/PGCN.py:37-51 turns into send/recv maps)
    {
        std::vector<unsigned long long> seen((size_t)n, 0ull);
        for (int64_t i = 0; i < n; ++i)
            for (int e = rowptr[(size_t)i]; e < rowptr[(size_t)i + 1]; ++e)
                seen[(size_t)colidx[(size_t)e]] |= 1ull << (part[(size_t)i] & 63);
        long long vol = 0;
        std::vector<long long> in((size_t)k, 0), rows((size_t)k, 0), ent((size_t)k, 0);
        for (int64_t j = 0; j < n; ++j) {
            const unsigned long long others = seen[(size_t)j] & ~(1ull << (part[(size_t)j] & 63));
            vol += __builtin_popcountll(others);
            for (int q = 0; q < k; ++q) if (others >> q & 1) ++in[(size_t)q];
            ++rows[(size_t)part[(size_t)j]];
            ent[(size_t)part[(size_t)j]] += rowptr[(size_t)j + 1] - rowptr[(size_t)j];
        }
        printf("halo rows total=%lld ; per part (rows, entries, halo in):", vol);
        for (int q = 0; q < k; ++q) printf(" (%lld,%lld,%lld)", rows[(size_t)q], ent[(size_t)q], in[(size_t)q]);
        printf("\n");
    }

    FILE* o = fopen(argv[4], "wb");
    if (!o) { perror(argv[4]); return 1; }
    std::vector<unsigned char> u8((size_t)n);
    for (int64_t i = 0; i < n; ++i) u8[(size_t)i] = (unsigned char)part[(size_t)i];
    fwrite(u8.data(), 1, (size_t)n, o);
    fclose(o);
    return 0;
}
