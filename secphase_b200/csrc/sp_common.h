// sp_common.h -- plain data shared by the CUDA kernels and their host launcher.
//
// The stage logic (sp_walk.cuh ... sp_score.cuh) is written as SP_HD functions over plain
// pointers.  In the product they are compiled by nvcc and only ever called from __global__
// kernels (sp_kernels.cu).  tests/hostsim/ compiles the very same headers with g++ to check the
// logic against the CPU oracle in the GPU-less build container; that harness is test-only and
// is not part of libsecphase_b200.so.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SP_HD __host__ __device__ __forceinline__
#define SP_D __device__ __forceinline__
#else
#define SP_HD inline
#define SP_D inline
#endif

// BAM CIGAR op codes (SAM spec; htslib BAM_C*)
enum { SP_CMATCH = 0, SP_CINS = 1, SP_CDEL = 2, SP_CREF_SKIP = 3, SP_CSOFT = 4, SP_CHARD = 5, SP_CPAD = 6,
       SP_CEQUAL = 7, SP_CDIFF = 8, SP_CSENTINEL = 15 };
enum { SP_FREVERSE = 16, SP_FSECONDARY = 256, SP_FSUPPLEMENTARY = 2048 };

// One refined operation of the CIGAR/cs walk (the state ptCigarIt exposes after each
// ptCigarIt_next, cigar_it.h:12-41), stored once per alignment so later stages never re-walk.
// Ends are implied by the next record (a sentinel closes the table):
//   sqe = next.sqs-1, rfe = next.rfs-1,
//   forward strand: rds_f = rdx,        rde_f = next.rdx-1
//   reverse strand: rde_f = rdx,        rds_f = next.rdx+1
struct SpOp {
    uint32_t oplen;  // op | len<<4   (len = iterator's len, cigar_it.h:19)
    int32_t sqs;
    int32_t rfs;
    int32_t rdx;
};

struct SpInitMarker {  // ptMarker_get_initial_markers output, ptMarker.c:42-75
    int32_t read_pos_f;
    int32_t base_idx;
    int32_t ref_pos;
    int32_t q;
};

struct SpEntry {  // one (position, alignment) marker of the filled list, ptMarker.h:34-41
    int32_t base_idx;
    int32_t ref_pos;
    int32_t q;      // base_q
    int32_t flags;  // bit0 is_match
};

struct SpBlock {  // ptBlock.h:23-38 (coordinates only)
    int32_t rfs, rfe, sqs, sqe, rds_f, rde_f;
};

struct SpItem {  // one probaln_glocal call of calc_local_baq, ptMarker.c:725-757
    int64_t ref_off;  // into the device reference codes
    int32_t aln;      // global alignment index (-1 in the stand-alone HMM batch API)
    int32_t blk;      // block index inside the alignment
    int32_t l_ref;
    int32_t l_query;
    int32_t q_sqs;    // first base of the window in the stored SEQ
    int32_t par_bw;   // conf.bw = |l_ref-l_query| + conf_b (ptMarker.c:754)
    int32_t row0;     // first marker-row slot
    int32_t n_rows;
    int64_t query_off;  // stand-alone API: offset into a byte-per-base query pool; pipeline: -1
    int64_t s_off;      // this instance's slice of the scaling-factor pool (l_query+2 doubles)
    int32_t op_first, op_last;  // refined ops the bq loop of ptMarker.c:767-785 examined for this window
};

struct SpRow {  // one query row whose MAP state/q is consumed (a marker inside an HMM window)
    int32_t item;
    int32_t t;         // 0-based row in the window
    int32_t entry;     // global entry index of the marker (pipeline), or -1
    int32_t expected;  // x + (t - y) of ptMarker.c:778, or INT32_MIN if no M/=/X op visited the base
    int32_t state;     // out
    int32_t q;         // out
    double pmax;       // out: normalised max posterior
};

// Scalars every kernel needs.
struct SpConst {
    int32_t baq_flag, consensus, indel_threshold, min_q, set_q, flank_margin;
    // --writeBam (secphase.c:182-189): the MAP state/q of EVERY row of the write-back range
    // [10, l_query-10) of each HMM window is needed, not just the marker rows (ptMarker.c:763-786)
    int32_t full_baq, pad0;
    double conf_b;
    // HMM constants derived on the host exactly as C evaluates them (float sub-expressions kept)
    double m0f;      // (double)(1 - d - d)   evaluated in float
    double omd_f;    // (double)(1 - d)       evaluated in float (bM numerator)
    double d_d;      // (double)d
    double ome_f;    // (double)(1 - e)       evaluated in float  (m[3] factor and m[6])
    double e_d;      // (double)e             (m[4] factor and m[8])
    float d_f, omd_ff;  // float d and float (1-d) for the float divisions bM, bI
    double em_match;    // 1. - (double)qual
    double em_mis;      // (double)qual * EM
    double qthr[102];   // qthr[n], n=1..101: largest t=1-max with (int)(-4.343*log(t)+.499) >= n
    double sc_match[256];  // -1 * reverse_quality(q)          ptMarker.c:298-304,315
    double sc_mis[256];    // -1 * q - 10 * log(3)             ptMarker.c:319
};

#define SP_INT_MIN (-2147483647 - 1)
#define SP_MAX_ALN_PER_GROUP_C 10  // secphase.c:286
#define SP_BLOCK_MARGIN 10    // ptMarker.c:697
#define SP_MAX_BLOCK_LEN 1000 // secphase.c:164

// group status bits
enum { SP_GERR_BLOCK_CAP = 1, SP_GERR_MARKER_CAP = 2, SP_GERR_OP_CAP = 4, SP_GERR_BADOP = 8 };

// Band classes of the HMM kernels: instances of one class share a launch (same shared-memory
// slab per warp).  Classes 0..SP_N_CLASSES-2 run the shared-memory-band kernel (sp_hmm2.cuh).
// The first SP_N_EXACT_CLASSES are exact band half-widths 20, 21, 22 (|Lr-Lq| = 0, 1, 2 with the
// presets' -b20: 87 % of the band cells of the HiFi workload) and get the fully unrolled row bodies;
// the other bounds are where one more warp's slab stops fitting an SM (5,4,3,2,1 warps of 2*bw+3
// cells x 768 B in 227 KB).  The last class is the generic kernel (sp_hmm.cuh), sized per launch.
// -DSP_N_EXACT_CLASSES=6 adds unrolled bodies for bw 23..25 (another 11 % of the cells): measured
// +3 % with three batches in flight but a longer single-batch launch set (three more small
// launches, each one poorly averaged round), so it is not the default (profiles/r01_exact_classes_v17.json).
#ifndef SP_N_EXACT_CLASSES
#define SP_N_EXACT_CLASSES 3
#endif
#define SP_N_CLASSES (SP_N_EXACT_CLASSES + 6)
SP_HD int sp_class_bw(int cls) {
    return cls < SP_N_EXACT_CLASSES ? 20 + cls
           : cls == SP_N_EXACT_CLASSES ? 27 : cls == SP_N_EXACT_CLASSES + 1 ? 36 : cls == SP_N_EXACT_CLASSES + 2 ? 48
           : cls == SP_N_EXACT_CLASSES + 3 ? 62 : cls == SP_N_EXACT_CLASSES + 4 ? 94 : 0;
}
SP_HD int sp_band_class(int bw) {
    return bw <= 20 ? 0 : bw <= 20 + SP_N_EXACT_CLASSES - 1 ? bw - 20
           : SP_N_EXACT_CLASSES + (bw <= 27 ? 0 : bw <= 36 ? 1 : bw <= 48 ? 2 : bw <= 62 ? 3 : bw <= 94 ? 4 : 5);
}
SP_HD int sp_class_unrolled_cells(int cls) {  // NC of sp_hmm2_instance
    return cls < SP_N_EXACT_CLASSES ? 2 * sp_class_bw(cls) + 1 : 0;
}

SP_HD int sp_min(int a, int b) { return a < b ? a : b; }
SP_HD int sp_max(int a, int b) { return b < a ? a : b; }
