/*
 * idcodec.h -- C ABI of the B200-native vector-id codec (libidcodec.so).
 *
 * This is the drop-in boundary for the ID-compression hot path of
 * facebookresearch/vector_db_id_compression. Every entry point names the
 * reference interface it replaces (paths relative to the reference tree).
 * Plain pointers and sizes only; no C++ or torch types. All functions return
 * IDC_OK (0) or a negative error code; idc_last_error() gives the message of
 * the last failure on the calling thread. No exceptions cross this boundary.
 *
 * There is NO CPU fallback: every codec entry point launches sm_100a CUDA
 * kernels and fails with IDC_ERR_CUDA when no device is usable.
 *
 * Vocabulary
 *   list   an IVF inverted list or an NSG adjacency row: a set of ids.
 *   unit   what one rANS stream covers. A list of <= max_unit ids is one unit;
 *          longer lists are cut into consecutive runs of max_unit id-sorted
 *          ids (the reference codec only round-trips sets of <= 65536 ids, see
 *          DESIGN.md), each an independent reference-compatible stream.
 *   blob   the device-resident compressed form of many lists.
 */
#ifndef IDCODEC_H
#define IDCODEC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IDC_OK 0
#define IDC_ERR_ARG (-1)      /* bad argument (null pointer, unsupported width, ...) */
#define IDC_ERR_CUDA (-2)     /* CUDA runtime failure / no device */
#define IDC_ERR_DOMAIN (-3)   /* input outside the codec's domain (id >= 2^32, unsorted with IDC_F_SORTED, ...) */
#define IDC_ERR_STREAM (-4)   /* a stream violated a decoder invariant (corrupt blob) */
#define IDC_ERR_NOMEM (-5)

/* where a data pointer lives */
#define IDC_MEM_HOST 0
#define IDC_MEM_DEVICE 1

/* encode flags */
#define IDC_F_SORTED 1u          /* ids of every list are already ascending (Faiss add order); verified on device */
#define IDC_F_PRECISION_SAFE 2u  /* precision = bit_length(max_id) instead of the reference's ceil(log2(max_id)) */
#define IDC_F_WANT_ORDER 4u      /* record the sample order (the permutation the reference applies to the codes) */

#define IDC_MAX_UNIT_DEFAULT 65536u

typedef struct idc_ctx idc_ctx;
typedef struct idc_roc_blob idc_roc_blob;
typedef struct idc_ef_blob idc_ef_blob;
typedef struct idc_bits_blob idc_bits_blob;

/* ------------------------------------------------------------------ context */

/* Bind to a CUDA device, create the stream the codec launches on and upload
 * the constant tables (mt19937(1234) words of codec.h:16-18,38; reciprocals
 * for the uniform pop/push of codec.cpp:21-63). No device-global state is
 * touched unless the environment asks for it: IDC_L2_FETCH=32|64|128 sets
 * cudaLimitMaxL2FetchGranularity for the lifetime of the context (the ROC
 * kernels read isolated 32-byte sectors); idc_ctx_destroy restores it. */
int idc_ctx_create(int device, idc_ctx** out);
/* Same, but launch on an existing cudaStream_t (e.g. torch's current stream). */
int idc_ctx_create_on_stream(int device, void* cuda_stream, idc_ctx** out);
/* Fails with IDC_ERR_ARG (and leaves the context intact) while blobs created by it are still alive: a blob returns
 * its arrays to the context's pool when it is freed. */
int idc_ctx_destroy(idc_ctx* ctx);
int idc_ctx_synchronize(idc_ctx* ctx);
/* number of kernels this context has launched so far (bench.py "gpu_launches") */
uint64_t idc_ctx_launch_count(const idc_ctx* ctx);
/* milliseconds spent in the kernels of the last encode/decode call, measured
 * with CUDA events on the codec stream (0 if timing is disabled) */
int idc_ctx_set_timing(idc_ctx* ctx, int enable);
float idc_ctx_last_kernel_ms(const idc_ctx* ctx);
/* per-kernel breakdown of the last call: up to cap (name, ms) pairs */
int idc_ctx_last_kernel_breakdown(const idc_ctx* ctx, const char** names, float* ms, int cap);

const char* idc_last_error(void);
int idc_version(void);

/* ------------------------------------------------------------------ ROC ----
 * Random Order Coding: bits-back rANS over a set.
 * Replaces, in bulk over all lists:
 *   compress()                       custom_invlist_cpp/codec.cpp:123-138
 *   CompressedIDInvertedListsFenwickTree ctor loop
 *                                    custom_invlist_cpp/custom_invlists_impl.cpp:147-194
 *   ROCNSGGraph ctor loop            alt-graph-index/altid_impl.cpp:108-149
 * The emitted (head, words) of every unit is bit-identical to ANSState{head,
 * stack} (codec.h:13-45) produced by the reference on the same id set and
 * precision.
 *
 * offsets: HOST array of nlist+1 element offsets (CSR) into ids.
 * ids:     int64 (id_bytes = 8, faiss::idx_t) or int32 (id_bytes = 4), host or device.
 *          IDC_MEM_HOST with IDC_F_SORTED: the upload is cut into chunks of whole units and
 *          overlapped with the kernels (a size class of units starts when its chunk is there);
 *          page-locked host memory makes the copies asynchronous. Input errors (an id >= 2^32,
 *          unsorted ids behind IDC_F_SORTED) are reported by the return code in every case.
 */
int idc_roc_encode(
        idc_ctx* ctx,
        uint64_t nlist,
        const uint64_t* offsets,
        const void* ids,
        int id_bytes,
        int ids_mem,
        uint32_t flags,
        uint32_t max_unit,
        idc_roc_blob** out);

/* NSG rows: data is N x K int32, each row terminated by the first -1
 * (altid_impl.cpp:110-117). Rows need not be sorted. */
int idc_roc_encode_rows(
        idc_ctx* ctx,
        uint64_t nrows,
        uint32_t K,
        const int32_t* data,
        int data_mem,
        uint32_t flags,
        idc_roc_blob** out);

typedef struct {
    uint64_t nlist;
    uint64_t nunits;
    uint64_t total_ids;
    uint64_t total_words;      /* 32-bit stream words over all units */
    uint64_t ans_bytes;        /* sum over non-empty units of ANSState::size() = 8 + 4*words (codec.h:42-44) */
    uint64_t device_bytes;     /* HBM held by the blob */
    uint32_t max_unit;
    uint32_t row_stride;       /* K for row blobs, 0 for CSR blobs */
} idc_roc_info;

int idc_roc_blob_info(const idc_roc_blob* blob, idc_roc_info* info);

/* Copy the blob's arrays to HOST memory (any pointer may be NULL to skip it).
 *   list_offsets[nlist+1]   ids per list (CSR)
 *   unit_offsets[nlist+1]   first unit of each list
 *   unit_n[nunits]          ids in the unit
 *   unit_precision[nunits]  id_symbol_precision (custom_invlists_impl.h:62)
 *   unit_heads[nunits]      ANSState::head
 *   word_offsets[nunits+1]  start of each unit's stack in words[] (bottom -> top)
 *   words[total_words]      ANSState::stack contents
 */
int idc_roc_blob_export(
        const idc_roc_blob* blob,
        uint64_t* list_offsets,
        uint64_t* unit_offsets,
        uint32_t* unit_n,
        uint8_t* unit_precision,
        uint64_t* unit_heads,
        uint64_t* word_offsets,
        uint32_t* words);

/* Build a blob from HOST arrays, e.g. ANS states produced by the reference
 * (one unit per list; unit_n[l] = list size). */
int idc_roc_blob_import(
        idc_ctx* ctx,
        uint64_t nlist,
        const uint32_t* unit_n,
        const uint8_t* unit_precision,
        const uint64_t* unit_heads,
        const uint64_t* word_offsets,
        const uint32_t* words,
        idc_roc_blob** out);

/* The wire form of a blob: what a gather over NCCL (or a file) carries. Per unit, in unit order: precision,
 * head, number of stream words, the decoder's id-range hints [lo, hi] (min / max id of the unit: they only steer
 * the decoder's speed, never its result); then all stream words back to back. `mem` says where the destination
 * buffers live (IDC_MEM_HOST or IDC_MEM_DEVICE: device -> device copies, nothing touches the host). Any pointer
 * may be NULL. Sizes: nunits entries each (idc_roc_blob_info), words: total_words. */
int idc_roc_blob_export_payload(
        const idc_roc_blob* blob,
        int mem,
        uint8_t* unit_precision,
        uint64_t* unit_heads,
        uint32_t* unit_nwords,
        uint32_t* unit_lo,
        uint32_t* unit_hi,
        uint32_t* words);

/* The inverse: build a blob over nlist lists (list_offsets: HOST CSR of the id counts; lists longer than max_unit
 * are split into units exactly as idc_roc_encode splits them) from per-unit payload arrays in unit order, HOST or
 * DEVICE per `mem`. Used on the rank that owns an index to re-assemble the blobs its peers encoded (the units of
 * consecutive ranks concatenate to the units of the whole index), and by the file loader. unit_lo / unit_hi may
 * be NULL (hints derived from the precision). The result is indistinguishable from the blob idc_roc_encode
 * returns for the whole index. */
int idc_roc_blob_assemble(
        idc_ctx* ctx,
        uint64_t nlist,
        const uint64_t* list_offsets,
        uint32_t max_unit,
        int mem,
        const uint8_t* unit_precision,
        const uint64_t* unit_heads,
        const uint32_t* unit_nwords,
        const uint32_t* unit_lo,
        const uint32_t* unit_hi,
        const uint32_t* words,
        uint64_t total_words,
        idc_roc_blob** out);

/* Sample order recorded with IDC_F_WANT_ORDER: order[(offsets[l] - offsets[0]) + t]
 * (rebased to the first list, like every decode output) is the
 * position inside list l (input order) of the id emitted at step t, which is
 * also index t of the decoded list -- the permutation applied to codes at
 * custom_invlists_impl.cpp:189-193. */
int idc_roc_blob_order(const idc_roc_blob* blob, uint32_t* order, int order_mem);

int idc_roc_blob_free(idc_roc_blob* blob);

/* Flat file form of a blob (layout: csrc/idc_file.h): the list CSR and the wire payload of
 * idc_roc_blob_export_payload. The reference keeps its ANS states in memory only (ans_states,
 * custom_invlists_impl.h:59; altid_impl.h:58); an index built once is saved and loaded with these instead of being
 * re-encoded. A loaded blob is indistinguishable from the saved one (row blobs included). */
int idc_roc_blob_save(const idc_roc_blob* blob, const char* path);
int idc_roc_blob_load(idc_ctx* ctx, const char* path, idc_roc_blob** out);

/* Replaces decompress() (codec.cpp:140-152) /
 * CompressedIDInvertedListsFenwickTree::get_ids (custom_invlists_impl.cpp:210-219)
 * in bulk. list_nos (HOST, may be NULL = all lists in order) selects nsel lists;
 * ids_out receives them concatenated, each list in the reference's decode
 * order (for multi-unit lists: unit after unit). out_offsets (HOST, nsel+1,
 * may be NULL) receives the CSR offsets of the output.
 * Whole-index decodes into HOST memory (list_nos == NULL, >= 2^22 ids) overlap the download with the kernels; when
 * ids_out is pinned or registered host memory (cudaHostAlloc / cudaHostRegister) the ids of the longest units leave
 * while their chains are still running. Pageable memory works, without that overlap. */
int idc_roc_decode(
        idc_ctx* ctx,
        const idc_roc_blob* blob,
        const uint64_t* list_nos,
        uint64_t nsel,
        void* ids_out,
        int id_bytes,
        int out_mem,
        uint64_t* out_offsets);

/* The id-translation step of search_IVF_defer_id_decoding
 * (custom_invlists_impl.cpp:464-525): labels hold (list_no << 32 | offset)
 * pairs as written by search_preassigned(store_pairs = true); `offset` counts
 * in the list's decode order (= the order of its re-laid codes). Every distinct
 * hit list is decoded ONCE, in one bulk launch (the reference groups the hits
 * by list and decodes the hit lists under OpenMP, :477-525), then the ids are
 * gathered on the device. Negative labels pass through unchanged. labels and
 * ids_out are n int64 values, HOST or DEVICE per their *_mem arguments. */
int idc_roc_translate(
        idc_ctx* ctx,
        const idc_roc_blob* blob,
        const int64_t* labels,
        int labels_mem,
        uint64_t n,
        int64_t* ids_out,
        int out_mem);

/* ROCNSGGraph::get_neighbors (altid_impl.cpp:153-165) for many rows at once:
 * out is nsel x K int32; entries past the row's length are set to -1;
 * counts (may be NULL) receives the true neighbour count per row (the
 * reference returns K). row_nos HOST or DEVICE per rows_mem;
 * NULL = all rows in order. A call with host row numbers and a host output of
 * at most 64 KB (one row, a row and its neighbours' rows) goes through the
 * context's pinned mailbox: one launch and one stream synchronisation, no
 * copies. The same holds for idc_ef_decode_rows, idc_ef_select and
 * idc_wt_select. */
int idc_roc_decode_rows(
        idc_ctx* ctx,
        const idc_roc_blob* blob,
        const int32_t* row_nos,
        int rows_mem,
        uint64_t nsel,
        int32_t* out,
        uint32_t* counts,
        int out_mem);

/* ------------------------------------------------------------ Elias-Fano ----
 * Replaces, in bulk:
 *   elias_fano_builder + push_back        elias_fano.hpp:22-57
 *   CompressedIDInvertedListsEliasFano ctor custom_invlists_impl.cpp:229-284
 *   EliasFanoNSGGraph ctor                altid_impl.cpp:53-90
 * Per list: l = msb(max_id / m), low bits m*l, high bits (m+1)+(max_id>>l)+1,
 * both LSB-first in 64-bit words (succinct bit_vector layout). A list whose
 * high-bits vector would reach 2^32 bits (more than ~1.4e9 ids) is rejected
 * with IDC_ERR_DOMAIN.
 * Device id buffers are read in whole 16-byte pieces (bulk copies): up to 8 / 12 bytes past the last id of the
 * array may be read (never used). Any cudaMalloc'd / framework-allocated buffer allows that (allocations are
 * 256-byte granular); host buffers are staged by the library.
 */
int idc_ef_encode(
        idc_ctx* ctx,
        uint64_t nlist,
        const uint64_t* offsets,
        const void* ids,
        int id_bytes,
        int ids_mem,
        uint32_t flags,
        idc_ef_blob** out);

int idc_ef_encode_rows(
        idc_ctx* ctx,
        uint64_t nrows,
        uint32_t K,
        const int32_t* data,
        int data_mem,
        uint32_t flags,
        idc_ef_blob** out);

typedef struct {
    uint64_t nlist;
    uint64_t total_ids;
    uint64_t low_words;        /* 64-bit words */
    uint64_t high_words;
    uint64_t bits_total;       /* sum over lists of m_low_bits.size() + m_high_bits.size() (custom_invlists_impl.cpp:277) */
    uint64_t device_bytes;
    uint32_t row_stride;
} idc_ef_info;

int idc_ef_blob_info(const idc_ef_blob* blob, idc_ef_info* info);

/* HOST export: list_offsets[nlist+1], l[nlist], universe[nlist] (max id),
 * low_offsets/high_offsets[nlist+1] in 64-bit words, low[], high[]. */
int idc_ef_blob_export(
        const idc_ef_blob* blob,
        uint64_t* list_offsets,
        uint8_t* l,
        uint64_t* universe,
        uint64_t* low_offsets,
        uint64_t* high_offsets,
        uint64_t* low,
        uint64_t* high);

/* The inverse of idc_ef_blob_export: list_offsets[nlist+1] (HOST CSR of the list lengths), universe[nlist] (HOST,
 * max id of each list = the n passed to the elias_fano builder, custom_invlists_impl.cpp:262-263), the two bit vectors
 * (HOST or DEVICE per `mem`; sizes and per-list offsets follow from (universe, length) as in elias_fano.hpp:28-29).
 * row_stride != 0 makes it a row blob (idc_ef_decode_rows). The select samples and the decoder's chunk directory are
 * rebuilt on the device; a bit vector that does not hold one set bit per id is refused (IDC_ERR_ARG). */
int idc_ef_blob_import(
        idc_ctx* ctx,
        uint64_t nlist,
        const uint64_t* list_offsets,
        const uint64_t* universe,
        uint32_t row_stride,
        const uint64_t* low,
        const uint64_t* high,
        int mem,
        idc_ef_blob** out);

/* Flat file form (csrc/idc_file.h): the list CSR, the universes and the two bit vectors (ef_bitstreams,
 * custom_invlists_impl.h:76; altid_impl.h:43). */
int idc_ef_blob_save(const idc_ef_blob* blob, const char* path);
int idc_ef_blob_load(idc_ctx* ctx, const char* path, idc_ef_blob** out);

int idc_ef_blob_free(idc_ef_blob* blob);

/* select_enumerator over whole lists: CompressedIDInvertedListsEliasFano::get_ids
 * (custom_invlists_impl.cpp:292-311), elias_fano.hpp:210-249. */
int idc_ef_decode(
        idc_ctx* ctx,
        const idc_ef_blob* blob,
        const uint64_t* list_nos,
        uint64_t nsel,
        void* ids_out,
        int id_bytes,
        int out_mem,
        uint64_t* out_offsets);

/* EliasFanoNSGGraph::get_neighbors (altid_impl.cpp:92-101), many rows. */
int idc_ef_decode_rows(
        idc_ctx* ctx,
        const idc_ef_blob* blob,
        const int32_t* row_nos,
        int rows_mem,
        uint64_t nsel,
        int32_t* out,
        uint32_t* counts,
        int out_mem);

/* elias_fano::select (elias_fano.hpp:141-145) /
 * CompressedIDInvertedListsEliasFano::get_single_id (custom_invlists_impl.cpp:314-318)
 * for nq (list_no, offset) pairs. */
int idc_ef_select(
        idc_ctx* ctx,
        const idc_ef_blob* blob,
        const uint64_t* list_nos,
        const uint64_t* offsets_in_list,
        uint64_t nq,
        int query_mem,
        int64_t* ids_out,
        int out_mem);

/* ---------------------------------------------------------- wavelet tree ----
 * Replaces CompressedIDInvertedListsWaveletTree
 * (custom_invlist_cpp/custom_invlists_impl.cpp:346-397, .h:100-124): ONE
 * structure over the sequence S[id] = list_no, id in [0, ntotal) (:354-362),
 * answering get_single_id(list_no, offset) = wt.select(offset + 1, list_no)
 * (:377-379) and get_ids = all offsets of a list (:381-392).
 *
 * The lists must partition [0, ntotal), ntotal = offsets[nlist] - offsets[0],
 * each list strictly ascending -- the reference's asserts (:358-359) -- else
 * IDC_ERR_DOMAIN. wt_type 0 = sdsl::wt_int<> with plain bit vectors; wt_type 1 =
 * sdsl::wt_int<rrr_vector<63>> (custom_invlists_impl.h:105, .cpp:371-372): the same
 * structure with every level stored as RRR(63) blocks -- class (ones among 63 bits)
 * + offset in the combinatorial number system, eight blocks and 8 verbatim bits per
 * 512-bit rank block (csrc/wt_core.cuh) -- selects decode one 63-bit block per level,
 * whole-list decodes expand the levels once.
 * SDSL is a third-party dependency absent from the reference tree, so the word
 * layout and size_in_bytes() of its wt_int / rrr_vector are not reproduced:
 * select values are exact, the structure is a wavelet matrix of
 * bit_length(nlist - 1) levels x ntotal bits plus rank / select directories.
 */
typedef struct idc_wt_blob idc_wt_blob;

int idc_wt_encode(
        idc_ctx* ctx,
        uint64_t nlist,
        const uint64_t* offsets,
        const void* ids,
        int id_bytes,
        int ids_mem,
        int wt_type,
        idc_wt_blob** out);

typedef struct {
    uint64_t nlist;
    uint64_t total_ids;
    uint64_t bits_bytes;       /* levels x ntotal bits, padded to 512-bit blocks; wt_type 1: classes + pointers + offset streams */
    uint64_t aux_bytes;        /* rank directory, select samples, list start table */
    uint64_t device_bytes;
    uint32_t levels;
    uint32_t wt_type;
} idc_wt_info;

int idc_wt_blob_info(const idc_wt_blob* blob, idc_wt_info* info);

/* HOST export (any pointer may be NULL): list_offsets[nlist+1];
 * bits[levels * words] with words = 8 * ceil(ntotal / 512), level after level, LSB first;
 * rank[levels * (words / 8 + 1)]: ones before each 512-bit block, last entry = ones of the level;
 * sel1 / sel0[levels * (ntotal / 2048 + 2)]: block of one / zero number m * 2048;
 * start[nlist]: position of the list's first id below the last level. */
int idc_wt_blob_export(
        const idc_wt_blob* blob,
        uint64_t* list_offsets,
        uint64_t* bits,
        uint32_t* rank,
        uint32_t* sel1,
        uint32_t* sel0,
        uint32_t* start);

/* The inverse of idc_wt_blob_export (arrays HOST or DEVICE per `mem`), and the flat file form of the index (the
 * reference's `wt` member, custom_invlists_impl.h:104-105; csrc/idc_file.h). */
int idc_wt_blob_import(
        idc_ctx* ctx,
        uint64_t nlist,
        const uint64_t* list_offsets,
        int wt_type,
        const uint64_t* bits,
        const uint32_t* rank,
        const uint32_t* sel1,
        const uint32_t* sel0,
        const uint32_t* start,
        int mem,
        idc_wt_blob** out);
/* wt_type = 1 only: the compressed arrays as they lie in HBM (HOST copies, any pointer may be NULL): cls[levels * nblk]
 * (eight 6-bit classes + the block's 8 tail bits per 512-bit block, nblk = ceil(ntotal / 512)), ptr[levels * (nblk + 1)]
 * (bit offset of the block's first offset field in its level's stream), off_base[levels + 1] (first 64-bit word of each
 * level's stream; off_base[levels] = words in use), off[off_base[levels]]. idc_wt_blob_export returns the PLAIN levels
 * for either type (that is what idc_wt_blob_import and the file form carry). */
int idc_wt_blob_export_rrr(const idc_wt_blob* blob, uint64_t* cls, uint32_t* ptr, uint64_t* off_base, uint64_t* off);
int idc_wt_blob_save(const idc_wt_blob* blob, const char* path);
int idc_wt_blob_load(idc_ctx* ctx, const char* path, idc_wt_blob** out);

int idc_wt_blob_free(idc_wt_blob* blob);

/* get_single_id (custom_invlists_impl.cpp:377-379) for nq (list_no, offset)
 * pairs; a pair outside the index yields -1. */
int idc_wt_select(
        idc_ctx* ctx,
        const idc_wt_blob* blob,
        const uint64_t* list_nos,
        const uint64_t* offsets_in_list,
        uint64_t nq,
        int query_mem,
        int64_t* ids_out,
        int out_mem);

/* get_ids (custom_invlists_impl.cpp:381-392) for nsel lists (list_nos HOST,
 * NULL = all lists in order); same output convention as idc_ef_decode. */
int idc_wt_decode(
        idc_ctx* ctx,
        const idc_wt_blob* blob,
        const uint64_t* list_nos,
        uint64_t nsel,
        void* ids_out,
        int id_bytes,
        int out_mem,
        uint64_t* out_offsets);

/* ------------------------------------------------------- fixed-width packing
 * CompressedIDInvertedListsPackedBits (custom_invlists_impl.cpp:62-118),
 * CompactBitNSGGraph (altid_impl.cpp:20-51): values LSB-first, `bits` each. */
int idc_bits_pack(
        idc_ctx* ctx,
        uint64_t n,
        const void* vals,
        int val_bytes,
        int vals_mem,
        int bits,
        uint8_t* out,
        uint64_t out_bytes,
        int out_mem);

int idc_bits_unpack(
        idc_ctx* ctx,
        uint64_t n,
        const uint8_t* code,
        uint64_t code_bytes,
        int code_mem,
        int bits,
        void* out,
        int val_bytes,
        int out_mem);

#ifdef __cplusplus
}
#endif
#endif /* IDCODEC_H */
