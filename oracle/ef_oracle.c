/* TEST INFRASTRUCTURE ONLY -- never linked into or called from the product.
 *
 * CPU restatement, in plain C, of the Elias-Fano coding used by the reference
 * (elias_fano.hpp, a modified copy of ot/succinct's elias_fano.hpp). The
 * reference header depends on ot/succinct's bit_vector.hpp / darray.hpp /
 * broadword.hpp, which are NOT vendored under /root/reference
 * (install-dependencies.sh:8 clones HEAD of github.com/ot/succinct, unpinned;
 * README.md:47 links commit 669eebbdcaa0562028a22cb7c877e512e4f1210b), so the
 * reference class cannot be compiled here.
 *
 * Parity status: VALUE-LEVEL PINNED, BIT-LEVEL UNPINNED. Pinned by formula and
 * by the reference's own tests: decoded ids == sorted input
 * (test_compressed_ivfs.py:74-79, test_altid.py:28-44), l = msb(max_id / m)
 * (elias_fano.hpp:28), bit counts that feed compressed_ids_size_in_bytes
 * (custom_invlists_impl.cpp:277, altid_impl.cpp:86). The bit ORDER inside the
 * two bit vectors follows succinct's published conventions (bit i of a vector
 * is bit i%64 of 64-bit word i/64; append_bits ORs the value in at the current
 * bit position, LSB first) -- no reference test asserts those bits.
 */

#include <stdint.h>
#include <string.h>

/* broadword::msb: index of the highest set bit (x != 0). */
static int msb64(uint64_t x) {
    int r = 0;
    while (x >>= 1)
        r++;
    return r;
}

/* elias_fano_builder ctor, elias_fano.hpp:22-33. n = universe bound as passed
 * by the callers (= max id of the list, custom_invlists_impl.cpp:262-263,
 * altid_impl.cpp:75-77), m = number of elements. */
void oracle_ef_params(uint64_t n, uint64_t m, uint32_t* l_out, uint64_t* low_bits_out, uint64_t* high_bits_out) {
    uint32_t l = (m && n / m) ? (uint32_t)msb64(n / m) : 0;
    *l_out = l;
    *low_bits_out = m * l; /* m_low_bits.size() after m push_backs (:40-42) */
    *high_bits_out = (m + 1) + (n >> l) + 1; /* :29 */
}

static uint64_t words_for(uint64_t bits) { return (bits + 63) / 64; }

/* push_back loop, elias_fano.hpp:35-46. ids must be ascending and <= n.
 * low / high must hold words_for(low_bits) / words_for(high_bits) words. */
int oracle_ef_encode(uint64_t n, uint64_t m, const uint64_t* ids, uint64_t* low, uint64_t* high) {
    uint32_t l;
    uint64_t lb, hb;
    oracle_ef_params(n, m, &l, &lb, &hb);
    memset(low, 0, words_for(lb) * 8);
    memset(high, 0, words_for(hb) * 8);
    uint64_t mask = l ? (((uint64_t)1 << l) - 1) : 0;
    uint64_t last = 0;
    for (uint64_t i = 0; i < m; i++) {
        uint64_t v = ids[i];
        if (v < last || v > n)
            return -1; /* assert at elias_fano.hpp:36 */
        last = v;
        if (l) {
            uint64_t pos = i * l, lowv = v & mask;
            low[pos >> 6] |= lowv << (pos & 63);
            if ((pos & 63) + l > 64)
                low[(pos >> 6) + 1] |= lowv >> (64 - (pos & 63));
        }
        uint64_t hp = (v >> l) + i; /* :43 */
        high[hp >> 6] |= (uint64_t)1 << (hp & 63);
    }
    return 0;
}

static uint64_t get_low(const uint64_t* low, uint64_t i, uint32_t l) {
    if (!l)
        return 0;
    uint64_t pos = i * l, w = low[pos >> 6] >> (pos & 63);
    if ((pos & 63) + l > 64)
        w |= low[(pos >> 6) + 1] << (64 - (pos & 63));
    return w & (((uint64_t)1 << l) - 1);
}

/* select_enumerator from position 0, elias_fano.hpp:210-249: the i-th set bit
 * of the high vector at position h gives ((h - i) << l) | low_i. */
void oracle_ef_decode(uint64_t m, uint32_t l, const uint64_t* low, const uint64_t* high, uint64_t* out) {
    uint64_t i = 0, w = 0;
    while (i < m) {
        uint64_t bits = high[w];
        while (bits && i < m) {
            uint64_t h = w * 64 + (uint64_t)__builtin_ctzll(bits);
            bits &= bits - 1;
            out[i] = ((h - i) << l) | get_low(low, i, l);
            i++;
        }
        w++;
    }
}

/* elias_fano::select(k), elias_fano.hpp:141-145. */
uint64_t oracle_ef_select(uint64_t k, uint32_t l, const uint64_t* low, const uint64_t* high) {
    uint64_t seen = 0, w = 0;
    for (;;) {
        uint64_t c = (uint64_t)__builtin_popcountll(high[w]);
        if (seen + c > k)
            break;
        seen += c;
        w++;
    }
    uint64_t bits = high[w];
    for (uint64_t r = k - seen; r; r--)
        bits &= bits - 1;
    uint64_t h = w * 64 + (uint64_t)__builtin_ctzll(bits);
    return ((h - k) << l) | get_low(low, k, l);
}

/* Bulk loops over lists / rows in a flat layout: per list l, `ids` is CSR via
 * offsets; low_off / high_off (nlist+1 each) are word offsets into the two
 * packed arrays. universe[l] is the bound passed to the builder (max id). */
int oracle_ef_encode_lists(
        uint64_t nlist,
        const uint64_t* offsets,
        const uint64_t* ids,
        const uint64_t* universe,
        const uint64_t* low_off,
        const uint64_t* high_off,
        uint64_t* low,
        uint64_t* high) {
    int bad = 0;
    for (uint64_t k = 0; k < nlist; k++) {
        uint64_t m = offsets[k + 1] - offsets[k];
        if (!m)
            continue;
        if (oracle_ef_encode(universe[k], m, ids + offsets[k], low + low_off[k], high + high_off[k]))
            bad = 1;
    }
    return bad ? -1 : 0;
}

void oracle_ef_decode_lists(
        uint64_t nlist,
        const uint64_t* offsets,
        const uint8_t* l,
        const uint64_t* low_off,
        const uint64_t* high_off,
        const uint64_t* low,
        const uint64_t* high,
        uint64_t* out) {
    for (uint64_t k = 0; k < nlist; k++) {
        uint64_t m = offsets[k + 1] - offsets[k];
        if (m)
            oracle_ef_decode(m, l[k], low + low_off[k], high + high_off[k], out + offsets[k]);
    }
}

/* The same loops over lists with OpenMP (what the reference's plugin constructor does:
 * `#pragma omp parallel for` over list_no, custom_invlists_impl.cpp:234) -- bench.py's Elias-Fano cpu_baseline.
 * nthreads <= 0: all host threads. */
#include <omp.h>

int oracle_ef_encode_lists_mt(
        uint64_t nlist,
        const uint64_t* offsets,
        const uint64_t* ids,
        const uint64_t* universe,
        const uint64_t* low_off,
        const uint64_t* high_off,
        uint64_t* low,
        uint64_t* high,
        int nthreads) {
    int bad = 0;
    if (nthreads <= 0)
        nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads) reduction(| : bad)
    for (int64_t k = 0; k < (int64_t)nlist; k++) {
        uint64_t m = offsets[k + 1] - offsets[k];
        if (!m)
            continue;
        if (oracle_ef_encode(universe[k], m, ids + offsets[k], low + low_off[k], high + high_off[k]))
            bad |= 1;
    }
    return bad ? -1 : 0;
}

void oracle_ef_decode_lists_mt(
        uint64_t nlist,
        const uint64_t* offsets,
        const uint8_t* l,
        const uint64_t* low_off,
        const uint64_t* high_off,
        const uint64_t* low,
        const uint64_t* high,
        uint64_t* out,
        int nthreads) {
    if (nthreads <= 0)
        nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
    for (int64_t k = 0; k < (int64_t)nlist; k++) {
        uint64_t m = offsets[k + 1] - offsets[k];
        if (m)
            oracle_ef_decode(m, l[k], low + low_off[k], high + high_off[k], out + offsets[k]);
    }
}
