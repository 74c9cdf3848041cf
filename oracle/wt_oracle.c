/* TEST INFRASTRUCTURE ONLY -- never linked into or called from the product.
 *
 * CPU restatement of the wavelet-tree id index of the reference:
 * CompressedIDInvertedListsWaveletTree (custom_invlists_impl.cpp:346-397).
 *   ctor      S[id] = list_no for every id of every list (:354-362), then
 *             sdsl::construct_im(wt, S) (:367-372)
 *   get_single_id(list_no, offset) = wt.select(offset + 1, list_no) (:377-379)
 *   get_ids(list_no) = get_single_id for offset 0 .. list_size-1 (:381-392)
 *
 * The wavelet tree itself is third-party: simongog/sdsl-lite, UNPINNED (the
 * reference README tells the user to clone and install it; no version, no
 * vendored copy under /root/reference). Its published contract is restated:
 * wt.select(j, c) = position of the j-th (1-based) occurrence of c in S.
 *
 * Parity status: VALUE-LEVEL pinned (every select is defined by S alone, and
 * the reference's tests check exactly that: test_compressed_ivfs.py:37-41,
 * 128-132); BIT-LEVEL and size_in_bytes() UNPINNED (SDSL absent). The bit
 * layout built here is this repository's wavelet matrix (csrc/wt_core.cuh),
 * restated with plain loops so the device arrays can be compared word by word.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define WT_BLOCK_LOG 9
#define WT_SAMPLE_LOG 11

uint32_t oracle_wt_levels(uint64_t nlist) {
    uint32_t l = 1;
    while (l < 32 && ((uint64_t)1 << l) < nlist)
        l++;
    return l;
}

/* S[id] = list_no; returns 0, or -1 when the lists do not partition [0, n) in
 * ascending order (the reference's asserts, :358-359) */
int oracle_wt_sequence(uint64_t nlist, const uint64_t* offsets, const int64_t* ids, uint32_t* S) {
    uint64_t n = offsets[nlist] - offsets[0];
    for (uint64_t i = 0; i < n; i++)
        S[i] = 0xffffffffu;
    for (uint64_t l = 0; l < nlist; l++) {
        int64_t prev = -1;
        for (uint64_t e = offsets[l]; e < offsets[l + 1]; e++) {
            int64_t id = ids[e];
            if (id <= prev || (uint64_t)id >= n || S[id] != 0xffffffffu)
                return -1;
            S[id] = (uint32_t)l;
            prev = id;
        }
    }
    return 0;
}

/* the definition: position of occurrence number k (0-based) of c in S; -1 if there is none */
int64_t oracle_wt_select_seq(uint64_t n, const uint32_t* S, uint32_t c, uint64_t k) {
    for (uint64_t i = 0; i < n; i++)
        if (S[i] == c && k-- == 0)
            return (int64_t)i;
    return -1;
}

/* Wavelet matrix of S with its directories, array shapes as in idc_wt_blob_export:
 * bits[levels][words], rank[levels][nblk+1], sel1/sel0[levels][(n >> 11) + 2], start[nlist]. */
void oracle_wt_build(uint64_t nlist, uint64_t n, const uint32_t* S, uint64_t* bits, uint32_t* rank, uint32_t* sel1,
                     uint32_t* sel0, uint32_t* start) {
    uint32_t levels = oracle_wt_levels(nlist);
    uint64_t nblk = (n + (1u << WT_BLOCK_LOG) - 1) >> WT_BLOCK_LOG;
    uint64_t words = nblk * 8, rstride = nblk + 1, sstride = (n >> WT_SAMPLE_LOG) + 2;
    uint32_t* cur = (uint32_t*)malloc((n + 1) * sizeof(uint32_t));
    uint32_t* nxt = (uint32_t*)malloc((n + 1) * sizeof(uint32_t));
    memcpy(cur, S, n * sizeof(uint32_t));
    memset(bits, 0, levels * words * 8);
    memset(sel1, 0, levels * sstride * 4);
    memset(sel0, 0, levels * sstride * 4);
    for (uint32_t lev = 0; lev < levels; lev++) {
        uint32_t shift = levels - 1 - lev;
        uint64_t* B = bits + lev * words;
        uint32_t* R = rank + lev * rstride;
        uint64_t ones = 0, zeros = 0;
        for (uint64_t i = 0; i < n; i++) {
            if ((i & ((1u << WT_BLOCK_LOG) - 1)) == 0)
                R[i >> WT_BLOCK_LOG] = (uint32_t)ones;
            uint32_t b = (cur[i] >> shift) & 1u;
            if (b) {
                B[i >> 6] |= (uint64_t)1 << (i & 63);
                if ((ones & ((1u << WT_SAMPLE_LOG) - 1)) == 0)
                    sel1[lev * sstride + (ones >> WT_SAMPLE_LOG)] = (uint32_t)(i >> WT_BLOCK_LOG);
                ones++;
            } else {
                if ((zeros & ((1u << WT_SAMPLE_LOG) - 1)) == 0)
                    sel0[lev * sstride + (zeros >> WT_SAMPLE_LOG)] = (uint32_t)(i >> WT_BLOCK_LOG);
                zeros++;
            }
        }
        R[nblk] = (uint32_t)ones;
        /* stable partition: zeros first */
        uint64_t z = 0, o = zeros;
        for (uint64_t i = 0; i < n; i++) {
            if ((cur[i] >> shift) & 1u)
                nxt[o++] = cur[i];
            else
                nxt[z++] = cur[i];
        }
        uint32_t* t = cur;
        cur = nxt;
        nxt = t;
    }
    /* below the last level every list is one run; its start = first position holding it */
    for (uint64_t l = 0; l < nlist; l++)
        start[l] = 0;
    {
        uint64_t* size = (uint64_t*)calloc(nlist + 1, sizeof(uint64_t));
        for (uint64_t i = 0; i < n; i++)
            size[S[i]]++;
        for (uint64_t i = 0; i < n;) {
            uint32_t c = cur[i];
            start[c] = (uint32_t)i;
            i += size[c];
        }
        /* empty lists: where their run would be (between their neighbours in bit-reversed order) */
        uint64_t acc = 0;
        uint64_t span = (uint64_t)1 << levels;
        for (uint64_t r = 0; r < span; r++) {
            uint64_t c = 0;
            for (uint32_t i = 0; i < levels; i++)
                c |= ((r >> i) & 1u) << (levels - 1 - i);
            if (c < nlist) {
                if (size[c] == 0)
                    start[c] = (uint32_t)acc;
                acc += size[c];
            }
        }
        free(size);
    }
    free(cur);
    free(nxt);
}

static uint64_t scan_select(const uint64_t* B, uint64_t n, uint32_t b, uint64_t k) {
    for (uint64_t i = 0; i < n; i++)
        if (((B[i >> 6] >> (i & 63)) & 1u) == b && k-- == 0)
            return i;
    return ~(uint64_t)0;
}

/* select through the levels by linear scans (uses bits, rank[..][nblk] and start only) */
int64_t oracle_wt_select(uint64_t nlist, uint64_t n, const uint64_t* bits, const uint32_t* rank, const uint32_t* start,
                         uint32_t c, uint64_t k) {
    uint32_t levels = oracle_wt_levels(nlist);
    uint64_t nblk = (n + (1u << WT_BLOCK_LOG) - 1) >> WT_BLOCK_LOG;
    uint64_t p = (uint64_t)start[c] + k;
    for (int lev = (int)levels - 1; lev >= 0; lev--) {
        uint32_t b = (c >> (levels - 1 - (uint32_t)lev)) & 1u;
        uint64_t ones = rank[(uint64_t)lev * (nblk + 1) + nblk];
        if (b)
            p -= n - ones;
        p = scan_select(bits + (uint64_t)lev * nblk * 8, n, b, p);
        if (p == ~(uint64_t)0)
            return -1;
    }
    return (int64_t)p;
}

/* ---- wt_type = 1: RRR(63) blocks (sdsl::rrr_vector<63> in the reference, custom_invlists_impl.h:105; SDSL absent:
 * the framing below is this repository's, restated here with plain loops to pin the GPU arrays word for word).
 * A 512-bit rank block = eight 63-bit blocks + 8 verbatim bits. A 63-bit block -> class k = number of ones and
 * offset = sum over its ones (positions c_1 < c_2 < ...) of C(c_i, i) (combinatorial number system), stored in
 * ceil(log2 C(63, k)) bits. Per level: cls[blk] = eight 6-bit classes | tail << 48; ptr[blk] = bit offset of the
 * block's first offset field in the level's stream; streams lie one after the other, each padded to whole words + 1. */
static uint64_t BIN[64][64];
static uint32_t WID[64];
static int rrr_ready = 0;
static void rrr_init(void) {
    if (rrr_ready) return;
    for (int n = 0; n < 64; n++)
        for (int k = 0; k < 64; k++) BIN[n][k] = k == 0 ? 1 : (n == 0 ? 0 : BIN[n - 1][k - 1] + BIN[n - 1][k]);
    for (int k = 0; k < 64; k++) {
        uint64_t m = BIN[63][k] - 1;
        uint32_t w = 0;
        while (m) { w++; m >>= 1; }
        WID[k] = w;
    }
    rrr_ready = 1;
}
static int bit_of(const uint64_t* words, uint64_t i) { return (int)((words[i >> 6] >> (i & 63)) & 1u); }

/* returns the number of 64-bit words of `off` in use (off_base[levels]); off must be zeroed and large enough
 * (levels * (nblk * 8 + 1) words always suffice) */
uint64_t oracle_rrr_encode(uint64_t levels, uint64_t nblk, const uint64_t* bits, uint64_t* cls, uint32_t* ptr,
                           uint64_t* off_base, uint64_t* off) {
    rrr_init();
    uint64_t base = 0;
    for (uint64_t lev = 0; lev < levels; lev++) {
        const uint64_t* B = bits + lev * nblk * 8;
        uint64_t p = 0;
        off_base[lev] = base;
        for (uint64_t blk = 0; blk < nblk; blk++) {
            uint64_t cw = 0;
            ptr[lev * (nblk + 1) + blk] = (uint32_t)p;
            for (int j = 0; j < 8; j++) {
                uint32_t k = 0;
                uint64_t o = 0;
                for (int c = 0; c < 63; c++)
                    if (bit_of(B, blk * 512 + 63 * (uint64_t)j + c)) o += BIN[c][++k];
                cw |= (uint64_t)k << (6 * j);
                for (uint32_t t = 0; t < WID[k]; t++, p++)
                    if ((o >> t) & 1u) off[base + (p >> 6)] |= 1ull << (p & 63);
            }
            for (int t = 0; t < 8; t++) cw |= (uint64_t)bit_of(B, blk * 512 + 504 + t) << (48 + t);
            cls[lev * nblk + blk] = cw;
        }
        ptr[lev * (nblk + 1) + nblk] = (uint32_t)p;
        base += (p + 63) / 64 + 1;
    }
    off_base[levels] = base;
    return base;
}

/* the inverse: plain bits (levels * nblk * 8 words, zeroed by the caller) from the compressed arrays */
void oracle_rrr_decode(uint64_t levels, uint64_t nblk, const uint64_t* cls, const uint32_t* ptr, const uint64_t* off_base,
                       const uint64_t* off, uint64_t* bits) {
    rrr_init();
    for (uint64_t lev = 0; lev < levels; lev++) {
        uint64_t* B = bits + lev * nblk * 8;
        const uint64_t* S = off + off_base[lev];
        for (uint64_t blk = 0; blk < nblk; blk++) {
            const uint64_t cw = cls[lev * nblk + blk];
            uint64_t p = ptr[lev * (nblk + 1) + blk];
            for (int j = 0; j < 8; j++) {
                const uint32_t k = (uint32_t)(cw >> (6 * j)) & 63u;
                uint64_t o = 0;
                for (uint32_t t = 0; t < WID[k]; t++, p++) o |= (uint64_t)bit_of(S, p) << t;
                int c = 62;
                for (uint32_t i = k; i >= 1; i--) {
                    while (BIN[c][i] > o) c--;
                    const uint64_t pos = blk * 512 + 63 * (uint64_t)j + (uint64_t)c;
                    B[pos >> 6] |= 1ull << (pos & 63);
                    o -= BIN[c][i];
                    c--;
                }
            }
            for (int t = 0; t < 8; t++)
                if ((cw >> (48 + t)) & 1u) {
                    const uint64_t pos = blk * 512 + 504 + (uint64_t)t;
                    B[pos >> 6] |= 1ull << (pos & 63);
                }
        }
    }
}
