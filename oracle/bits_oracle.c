/* TEST INFRASTRUCTURE ONLY -- never linked into or called from the product.
 *
 * CPU restatement of the fixed-width id packing used by the reference's
 * "packed bits" baselines: CompressedIDInvertedListsPackedBits
 * (custom_invlists_impl.cpp:62-118) and CompactBitNSGGraph
 * (altid_impl.cpp:20-51). Both write through faiss::BitstringWriter /
 * BitstringReader [third-party, faiss/utils/hamming.h, not under
 * /root/reference]: values are appended LSB-first into a byte string, value k
 * occupying bits [k*bits, (k+1)*bits). BitstringReader_get_bits
 * (custom_invlists_impl.cpp:35-58) is the random-access read of the same
 * layout.
 *
 * Parity status: value-level pinned (round trip, reference tests
 * test_compressed_ivfs.py:28-29,134-135, test_altid.py:19-20); the byte layout
 * follows the published Faiss BitstringWriter convention.
 */
#include <stdint.h>
#include <string.h>

/* bits such that (1 << bits) >= ntotal + 1  (custom_invlists_impl.cpp:66-68,
 * altid_impl.cpp:22-23) */
int oracle_bits_for(uint64_t ntotal) {
    int bits = 0;
    while (((uint64_t)1 << bits) < ntotal + 1)
        bits++;
    return bits;
}

void oracle_bits_pack(uint64_t n, const uint64_t* vals, int bits, uint8_t* out, uint64_t out_bytes) {
    memset(out, 0, out_bytes);
    for (uint64_t k = 0; k < n; k++) {
        uint64_t pos = k * (uint64_t)bits;
        for (int b = 0; b < bits; b++, pos++)
            if ((vals[k] >> b) & 1)
                out[pos >> 3] |= (uint8_t)(1u << (pos & 7));
    }
}

uint64_t oracle_bits_get(const uint8_t* code, uint64_t k, int bits) {
    uint64_t pos = k * (uint64_t)bits, v = 0;
    for (int b = 0; b < bits; b++, pos++)
        v |= (uint64_t)((code[pos >> 3] >> (pos & 7)) & 1) << b;
    return v;
}

void oracle_bits_unpack(uint64_t n, const uint8_t* code, int bits, uint64_t* out) {
    for (uint64_t k = 0; k < n; k++)
        out[k] = oracle_bits_get(code, k, bits);
}
