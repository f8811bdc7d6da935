// mm_types.cuh — shared types of the stage-4 device pipeline (product code).
//
// Everything in mm_*.cuh is TELR_HD (__host__ __device__) sequential control logic that one lane
// of a warp executes between the warp-parallel primitives (chain inner loop, DP).  The same headers
// compile for the host so that tests/emu can drive the identical logic on a CPU without a GPU; the
// shipped library has no CPU execution path.
//
// Behavioural contract: `minimap2 -a -x <preset> contig reads` as invoked by the reference at
// TELR_te.py:503-506 (realignment), minimap2 2.22 per envs/telr.yml:45.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define TELR_HD __host__ __device__ __forceinline__
#define TELR_HDN __host__ __device__ __noinline__
#else
#define TELR_HD inline
#define TELR_HDN
#endif

// fp32 / fp64 arithmetic without fused multiply-add, identical on the device and in host builds (-ffp-contract=off)
#if defined(__CUDA_ARCH__)
#define TELR_FMUL(a, b) __fmul_rn((a), (b))
#define TELR_FADD(a, b) __fadd_rn((a), (b))
#define TELR_DMUL(a, b) __dmul_rn((a), (b))
#define TELR_DADD(a, b) __dadd_rn((a), (b))
#else   // host builds use -ffp-contract=off
#define TELR_FMUL(a, b) ((a) * (b))
#define TELR_FADD(a, b) ((a) + (b))
#define TELR_DMUL(a, b) ((a) * (b))
#define TELR_DADD(a, b) ((a) + (b))
#endif


namespace telr {

struct Anchor { uint64_t x, y; };

constexpr uint64_t SEED_LONG_JOIN = 1ULL << 40;
constexpr uint64_t SEED_IGNORE = 1ULL << 41;
constexpr uint64_t SEED_TANDEM = 1ULL << 42;
constexpr int PARENT_UNSET = -1;
constexpr int PARENT_TMP_PRI = -2;
constexpr int KSW_NEG_INF = -0x40000000;

constexpr int KSW_EXTZ_ONLY = 0x40;
constexpr int KSW_RIGHT = 0x02;
constexpr int KSW_REV_CIGAR = 0x80;
constexpr int KSW_APPROX_MAX = 0x08;

// minimap2 option state for the three presets (options.c), resolved on the host.
struct Opt {
    int k, w, hpc;
    int a, b, q, e, q2, e2, sc_ambi;
    int zdrop, zdrop_inv, end_bonus;
    int min_dp_max, min_ksw_len;
    int bw, bw_long, max_gap;
    int max_chain_skip, max_chain_iter, min_cnt, min_chain_score;
    int rmq_inner_dist, rmq_size_cap, rmq_rescue_size;
    float rmq_rescue_ratio, chn_pen_gap, chn_pen_skip;
    float mask_level;
    int mask_len;
    float pri_ratio;
    int best_n;
    float q_occ_frac, mid_occ_frac;
    int min_mid_occ, max_mid_occ;
    int max_max_occ, occ_dist;     // mm_seed_select: cap on occurrences, one rescued seed per occ_dist query bases
    uint32_t seed_term;      // Wang hash of opt->seed (map.c mm_map_frag)
    long long max_sw_mat;
    int rank_min_len;
    float rank_frac, max_clip_ratio;
};

// one chain / alignment region (minimap2 mm_reg1_t + the parts of mm_extra_t that matter here)
struct Reg {
    int32_t id, cnt, score, qs, qe, rs, re, parent, subsc, as, mlen, blen, n_sub, score0;
    uint32_t hash;
    uint8_t rev, inv, sam_pri, split, split_inv, strand_retained, has_p, need_fin;
    int32_t dp_score, dp_max, dp_max2, n_ambi, n_cigar;
    uint32_t cig;            // offset of this region's CIGAR in the problem's cigar arena
    int32_t fin_q, fin_t;    // where the joined CIGAR starts on query/target (deferred reg_finish)
    int32_t mapq;
};

// result of one DP call (ksw_extz_t)
struct DpRes {
    int32_t max, max_q, max_t, mqe, mqe_t, mte, mte_q, score, zdropped, reach_end;
    int32_t n_cigar;
    uint32_t *cigar;         // in the warp's scratch
    int32_t ll_score, ll_qe, ll_te;   // local-alignment probe (ksw_ll)
};

// one DP request emitted by the alignment state machine
struct DpTask {
    int32_t kind;            // 0 = two-piece affine extension/global DP, 1 = local score probe
    const uint8_t *q, *t;    // element i of the query is q[i * qstep] (complemented when qcomp)
    int32_t qstep, tstep, qcomp;
    int32_t qlen, tlen, w, zdrop, end_bonus, flag;
};

TELR_HD uint32_t wang_hash32(uint32_t key)
{
    key += ~(key << 15);
    key ^= (key >> 10);
    key += (key << 3);
    key ^= (key >> 6);
    key += ~(key << 11);
    key ^= (key >> 16);
    return key;
}

TELR_HD uint64_t mix64(uint64_t key)
{
    key = ~key + (key << 21);
    key = key ^ key >> 24;
    key = (key + (key << 3)) + (key << 8);
    key = key ^ key >> 14;
    key = (key + (key << 2)) + (key << 4);
    key = key ^ key >> 28;
    key = key + (key << 31);
    return key;
}

TELR_HD uint64_t mix64_masked(uint64_t key, uint64_t mask)
{
    key = (~key + (key << 21)) & mask;
    key = key ^ key >> 24;
    key = ((key + (key << 3)) + (key << 8)) & mask;
    key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4)) & mask;
    key = key ^ key >> 28;
    key = (key + (key << 31)) & mask;
    return key;
}

template <class T> TELR_HD T tmin(T a, T b) { return a < b ? a : b; }
template <class T> TELR_HD T tmax(T a, T b) { return a > b ? a : b; }

}  // namespace telr
