// Attribute gates on the GPU: multi-pattern substring match over the candidates' product text.
//
// Reference: calculate_gate_factor utils.py:88-101 (= _gate_factor app/app_product_search.py:228-236),
// applied to `agg_text[:6000]` of every pool member (app/app_product_search.py:297-302, app/test.py:291-297):
//
//     factor = 1.0;  for group in groups:  if not any(syn in text.lower() for syn in group): factor *= penalty
//
// Text store (built once at load, host side): for every product row the UTF-8 bytes of
// `str(agg_text)[:6000].lower()`, every text starting on a 16-byte boundary.  Patterns are the group
// members (ASCII: query tokens and the fixed COLORS / SYNONYMS sets), so a byte-wise search over UTF-8 is
// exact.  The 19 fixed groups are resolved from a per-row bitmap computed once by the same kernel
// (doc mode); only the query's free tokens (len >= 4, utils.py:79-80) scan text at query time.
//
// One warp per (query, candidate).  The text is staged through shared memory in 2 KB chunks with 16-byte
// loads (+ an overlap so that a match may straddle a chunk edge); every lane tests a strided set of start
// positions against a pattern, a warp vote ends a group at its first hit.
#include "rr_internal.h"
#include "rr_kernels.h"

namespace {

constexpr int GATE_WARPS = 8;
constexpr int GATE_CHUNK = 2048;
constexpr int GATE_OVERLAP = 64;                    // longest pattern matched inside a staged chunk is OVERLAP+1
constexpr int GATE_SLAB = GATE_CHUNK + GATE_OVERLAP + 16;

struct GateArgs {
    const unsigned char* text;        // byte blob
    const long long* text_off;        // [n_docs+1] byte offsets (16-byte aligned starts); text r = [off[r], off[r]+len[r])
    const int* text_len;              // [n_docs] byte lengths
    long long n_docs;
    const unsigned int* fixed_bits;   // [n_docs] bit g = fixed group g matches the row's text, or NULL
    const unsigned char* pat;         // pattern bytes
    const int* pat_off;               // [n_pat+1]
    const int* group_pat_off;         // [n_groups+1] patterns of a group
    const int* group_fixed;           // [n_groups] fixed-group id (bit of fixed_bits) or -1; may be NULL
    const int* query_group_off;       // [B+1] groups of a query (query mode)
    const long long* cand;            // [B, pool] product rows (query mode) or NULL (doc mode)
    int pool, B;
    int n_groups_doc;                 // doc mode: groups [0, n_groups_doc) are tested for every row
    double penalty;
    float* gate;                      // query mode: [B, pool]
    int* hits;                        // query mode, optional: [B, pool] number of matching groups
    unsigned int* out_bits;           // doc mode: [n_docs]
};

// does pattern p (global memory, m bytes) occur in text [t, t+len) ?  whole-warp call, uniform result
__device__ bool warp_find_global(const unsigned char* t, int len, const unsigned char* p, int m, int lane) {
    if (m == 0) return true;
    bool any = false;
    for (int base = 0; base + m <= len && !any; base += 32) {
        const int i = base + lane;
        bool ok = i + m <= len;
        for (int j = 0; ok && j < m; ++j) ok = t[i + j] == p[j];
        any = __any_sync(0xffffffffu, ok);
    }
    return any;
}

template <bool DOC_MODE>
__global__ void __launch_bounds__(GATE_WARPS * 32)
gate_kernel(GateArgs a) {
    __shared__ __align__(16) unsigned char s_text[GATE_WARPS][GATE_SLAB];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * GATE_WARPS + wib;
    const long long n_items = DOC_MODE ? a.n_docs : (long long)a.B * a.pool;
    if (gw >= n_items) return;

    long long r;
    int g0, g1;
    if constexpr (DOC_MODE) {
        r = gw; g0 = 0; g1 = a.n_groups_doc;
    } else {
        const int q = (int)(gw / a.pool);
        r = a.cand[gw];
        g0 = a.query_group_off[q]; g1 = a.query_group_off[q + 1];
    }
    const int n_g = min(g1 - g0, 32);
    if (r < 0 || r >= a.n_docs) {
        if constexpr (!DOC_MODE) { if (lane == 0) { a.gate[gw] = 1.0f; if (a.hits) a.hits[gw] = 0; } }
        return;
    }
    const unsigned char* text = a.text + a.text_off[r];
    const int len = a.text_len[r];

    unsigned int matched = 0u;                    // bit i = group g0+i has a hit
    unsigned int pending = (n_g >= 32) ? 0xffffffffu : ((1u << n_g) - 1u);
    // fixed groups: one bit test
    if (!DOC_MODE && a.fixed_bits != nullptr && a.group_fixed != nullptr) {
        const unsigned int bits = a.fixed_bits[r];
        for (int i = 0; i < n_g; ++i) {
            const int f = a.group_fixed[g0 + i];
            if (f >= 0) { pending &= ~(1u << i); if ((bits >> f) & 1u) matched |= 1u << i; }
        }
    }
    // patterns too long for the staged search, and empty patterns ("" in s is True)
    for (int i = 0; i < n_g; ++i) {
        if (!((pending >> i) & 1u)) continue;
        for (int p = a.group_pat_off[g0 + i]; p < a.group_pat_off[g0 + i + 1]; ++p) {
            const int m = a.pat_off[p + 1] - a.pat_off[p];
            if (m == 0 || m > GATE_OVERLAP + 1) {
                if (warp_find_global(text, len, a.pat + a.pat_off[p], m, lane)) { matched |= 1u << i; pending &= ~(1u << i); break; }
            }
        }
    }
    // staged search
    unsigned char* s = s_text[wib];
    for (int c = 0; c < len && pending != 0u; c += GATE_CHUNK) {
        const int avail = min(len - c, GATE_CHUNK + GATE_OVERLAP);
        __syncwarp();
        for (int v = lane; v * 16 < avail; v += 32)          // text starts are 16-byte aligned, c is a multiple of 16
            *reinterpret_cast<uint4*>(s + v * 16) = *reinterpret_cast<const uint4*>(text + c + v * 16);
        __syncwarp();
        for (int i = 0; i < n_g; ++i) {
            if (!((pending >> i) & 1u)) continue;
            bool hit = false;
            for (int p = a.group_pat_off[g0 + i]; p < a.group_pat_off[g0 + i + 1] && !hit; ++p) {
                const int m = a.pat_off[p + 1] - a.pat_off[p];
                if (m == 0 || m > GATE_OVERLAP + 1) continue;
                const unsigned char* pp = a.pat + a.pat_off[p];
                const unsigned char p0 = pp[0];
                const int n_start = min(GATE_CHUNK, avail - m + 1);         // start positions inside this chunk
                bool ok = false;
                for (int base = 0; base < n_start; base += 32) {
                    const int pos = base + lane;
                    bool mine = pos < n_start && s[pos] == p0;
                    for (int j = 1; mine && j < m; ++j) mine = s[pos + j] == pp[j];
                    if (__any_sync(0xffffffffu, mine)) { ok = true; break; }
                }
                hit = ok;
            }
            if (hit) { matched |= 1u << i; pending &= ~(1u << i); }
        }
    }
    if (lane == 0) {
        if constexpr (DOC_MODE) {
            a.out_bits[r] = matched;
        } else {
            double f = 1.0;
            for (int i = 0; i < n_g; ++i) if (!((matched >> i) & 1u)) f = __dmul_rn(f, a.penalty);   // `factor *= penalty`
            a.gate[gw] = (float)f;                                                                  // np.array(.., float32)
            if (a.hits) a.hits[gw] = __popc(matched);
        }
    }
}

}  // namespace

int rr_launch_gate_query(const uint8_t* d_text, const int64_t* d_text_off, const int32_t* d_text_len, int64_t n_docs,
                         const uint32_t* d_fixed_bits, const uint8_t* d_pat, const int32_t* d_pat_off,
                         const int32_t* d_group_pat_off, const int32_t* d_group_fixed, const int32_t* d_query_group_off,
                         int B, const int64_t* d_cand, int pool, double penalty, float* d_gate, int32_t* d_hits,
                         cudaStream_t stream) {
    const long long items = (long long)B * pool;
    if (items <= 0) return RR_OK;
    GateArgs a{};
    a.text = d_text; a.text_off = reinterpret_cast<const long long*>(d_text_off); a.text_len = d_text_len; a.n_docs = n_docs;
    a.fixed_bits = d_fixed_bits; a.pat = d_pat; a.pat_off = d_pat_off; a.group_pat_off = d_group_pat_off;
    a.group_fixed = d_group_fixed; a.query_group_off = d_query_group_off;
    a.cand = reinterpret_cast<const long long*>(d_cand); a.pool = pool; a.B = B; a.penalty = penalty;
    a.gate = d_gate; a.hits = d_hits;
    RrProfScope prof(RR_PROF_MISC, stream);
    gate_kernel<false><<<(unsigned)((items + GATE_WARPS - 1) / GATE_WARPS), GATE_WARPS * 32, 0, stream>>>(a);
    RR_LAUNCH_CHECK();
    return RR_OK;
}

int rr_launch_gate_bitmaps(const uint8_t* d_text, const int64_t* d_text_off, const int32_t* d_text_len, int64_t n_docs,
                           const uint8_t* d_pat, const int32_t* d_pat_off, const int32_t* d_group_pat_off, int n_groups,
                           uint32_t* d_bits, cudaStream_t stream) {
    if (n_docs <= 0) return RR_OK;
    GateArgs a{};
    a.text = d_text; a.text_off = reinterpret_cast<const long long*>(d_text_off); a.text_len = d_text_len; a.n_docs = n_docs;
    a.pat = d_pat; a.pat_off = d_pat_off; a.group_pat_off = d_group_pat_off; a.n_groups_doc = n_groups; a.out_bits = d_bits;
    RrProfScope prof(RR_PROF_MISC, stream);
    gate_kernel<true><<<(unsigned)((n_docs + GATE_WARPS - 1) / GATE_WARPS), GATE_WARPS * 32, 0, stream>>>(a);
    RR_LAUNCH_CHECK();
    return RR_OK;
}
