/*
 * haptools_cuda.h -- C ABI of libhaptools_cuda.so (B200 / sm_100a)
 *
 * The drop-in boundary for ONE hot path of CAST-genomics/haptools: the haplotype
 * transform.  The reference has no FFI seam for this path (it is ~75 lines of numpy);
 * the entry points below are what a ctypes binding inside the reference's four
 * transform methods binds instead of the numpy body:
 *
 *   Haplotypes.transform          haptools/data/haplotypes.py:1339-1413
 *   Haplotype.transform           haptools/data/haplotypes.py:463-511
 *   HaplotypesAncestry.transform  haptools/transform.py:91-149
 *   HaplotypeAncestry.transform   haptools/transform.py:31-68
 *
 * Conventions
 *   - Plain C: pointers + sizes, no C++/torch types.  All data pointers are DEVICE
 *     pointers unless the parameter name starts with `h_`.  The caller owns every
 *     buffer (the Python host layer allocates them as torch tensors); the library keeps
 *     no state between calls except the thread-local last-error string.
 *   - Every launcher is asynchronous on `stream` (a cudaStream_t passed as void*; NULL =
 *     the legacy default stream) and thread-safe.
 *   - Return value: HT_OK (0) or an HT_ERR_* code; ht_last_error() then describes it.
 *   - Strides are in BYTES.
 *
 * Data model
 *   key      one distinct (variant column, allele index[, ancestry label]) triple that at
 *            least one haplotype references; the reference's `alleles` dict entry
 *            (haplotypes.py:1375-1386).  The planner numbers keys so that keys of the
 *            same variant are consecutive (CSR `var_key_off`), which lets the pack kernel
 *            read each genotype column exactly once.
 *   plane    the bit vector over samples of one key on one chromosome copy (strand):
 *            bit = (GT[s, col, strand] == allele) [& (ANC[s, col, strand] == label)],
 *            i.e. one column of the reference's `equality_arr` (haplotypes.py:1404,
 *            transform.py:137,146) stored as 1 bit instead of 1 byte.
 *   layout   planes are tile-major:  planes[tile][key][word 0..31][strand 0..1]  (uint32),
 *            a tile being HT_TILE_SAMPLES = 1024 consecutive samples; bit b of word w of
 *            tile t is sample t*1024 + w*32 + b.  Bits of samples >= n_samples are 0.
 *            One (tile, key) record is 256 contiguous bytes.
 */
#ifndef HAPTOOLS_CUDA_H
#define HAPTOOLS_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HT_ABI_VERSION 9

#define HT_TILE_SAMPLES 1024 /* samples per plane tile            */
#define HT_TILE_WORDS 32     /* 32-bit words per tile per strand  */
#define HT_RECORD_WORDS 64   /* uint32 per (tile, key) record     */

#define HT_OK 0
#define HT_ERR_BAD_ARG 1     /* NULL pointer, negative size, misaligned workspace ... */
#define HT_ERR_CUDA 2        /* a CUDA runtime call or launch failed                  */
#define HT_ERR_UNSUPPORTED 3 /* layout / size outside what the kernels handle         */

/* bits of the optional `flags` word written by the pack kernels (QC side channel:
 * Genotypes.check_missing, haptools/data/genotypes.py:444-485) */
#define HT_FLAG_MISSING 1u      /* some referenced GT value >= 254                     */
#define HT_FLAG_NON_BIALLELIC 2u /* some referenced GT value in 2..253                 */

/* which pack kernel ht_pack_u8 should use; HT_PACK_AUTO picks from the strides */
#define HT_PACK_AUTO 0
#define HT_PACK_GENERIC 1      /* any strides: byte loads + __ballot_sync             */
#define HT_PACK_VARIANT_MAJOR 2 /* samples contiguous along a variant row (VCF reader
                                   layout, genotypes.py:183-186,210): 128-bit loads   */
#define HT_PACK_SAMPLE_MAJOR 3  /* variants contiguous along a sample row (PGEN reader
                                   layout, genotypes.py:1346-1352): 128-bit loads +
                                   bytewise bit accumulation down the sample rows     */
#define HT_PACK_SAMPLE_MAJOR_BIALLELIC 4 /* same layout, 16-byte aligned rows, tables->col_key01
                                   (with ancestry: col_key_range) given: one pass over the rows
                                   builds the planes of every allele-0/1 key with 128-byte
                                   coalesced record stores; variants listed in
                                   tables->var_other go through the generic kernel.  What
                                   HT_PACK_AUTO picks when it can.  The plain kernel runs with
                                   the slab mapping (CTA = 128 rows x 512 columns, 32-byte
                                   record stores; HT_PACK_SM2_KERNEL=tile: 1,024 rows x 64
                                   columns, 128-byte record stores).  With ancestry the
                                   line-consuming form runs (a warp reads 4 rows x 128 bytes,
                                   32-byte record stores); HT_PACK_ANC_KERNEL=warp in the
                                   environment selects the earlier form (measurements, tests) */

/* direction of ht_copy2d */
#define HT_COPY_H2D 1
#define HT_COPY_D2H 2
#define HT_COPY_D2D 3

/* ABI version of the loaded library (== HT_ABI_VERSION of the header it was built with). */
int ht_version(void);

/* Message for the last non-zero return on the calling thread ("" if none). */
const char* ht_last_error(void);

/* Number of plane tiles for n_samples, and bytes of the plane workspace for n_keys keys. */
int64_t ht_num_tiles(int64_t n_samples);
size_t ht_plane_bytes(int64_t n_samples, int64_t n_keys);

/*
 * Index tables of the pack kernels (all DEVICE pointers; built by the host planner,
 * haptools_b200/plan.py).
 *   var_col[V]        column (index along the variant axis of gt/anc) of each distinct
 *                     referenced variant, ascending.
 *   var_key_off[V+1]  CSR: keys var_key_off[v] .. var_key_off[v+1]-1 belong to var_col[v].
 *   key_allele[A]     allele index each key matches against (`alleles.index(allele)`,
 *                     haplotypes.py:1395; `int(allele != REF)`, transform.py:130).
 *   key_label[A]      ancestry label of each key (NULL iff no ancestry array is packed).
 *   col_var[n_cols]   optional dense inverse of var_col (-1 = column not referenced); needed
 *                     by HT_PACK_SAMPLE_MAJOR (it walks whole rows of gt).
 *   col_key01[2*n_cols]  optional: {key of allele 0, key of allele 1} of each column, -1 =
 *                     no such key; (-1, -1) for unreferenced columns and for the columns of
 *                     the variants in var_other.  Needed by HT_PACK_SAMPLE_MAJOR_BIALLELIC.
 *   col_key_range[2*n_cols]  optional, ancestry plans: {first key, key count} of each column
 *                     whose keys all have allele <= 1 and label < 8; {0, 0} otherwise.
 *   var_other[n_var_other]  indices (into var_col) of the variants NOT covered by col_key01 /
 *                     col_key_range (multi-allelic index, label >= 8); they are packed by
 *                     the generic kernel.
 */
typedef struct ht_pack_tables {
    int64_t n_cols;
    int64_t n_vars;
    int64_t n_keys;
    const uint32_t* var_col;
    const uint32_t* var_key_off;
    const uint8_t* key_allele;
    const uint8_t* key_label;
    const int32_t* col_var;
    const int32_t* col_key01;
    const int32_t* col_key_range;
    const uint32_t* var_other;
    int64_t n_var_other;
} ht_pack_tables;

/*
 * K1 -- pack.  Replaces the column gather `gts.subset(variants=...)`
 * (haplotypes.py:1388 -> genotypes.py:440) and `np.equal(allele_arr, gts.data[:, :, :2])`
 * (haplotypes.py:1404; transform.py:137) and, when `anc` is given,
 * `gts.ancestry[:, idxs] == ancestries[i]` (transform.py:146).
 *
 *   gt            genotype bytes (uint8 or numpy bool), element (s, v, c) at
 *                 gt + s*gt_s_samp + v*gt_s_var + c*gt_s_strand, c in {0,1}; never copied.
 *   anc           local-ancestry label bytes with its own strides, or NULL.
 *   tables        see ht_pack_tables (the struct itself is HOST memory).
 *   planes        workspace of ht_plane_bytes(n_samples, A) bytes, 16-byte aligned;
 *                 fully overwritten.
 *   flags         optional (may be NULL) uint32 that the kernel ORs HT_FLAG_* bits into.
 *   mode          HT_PACK_*.
 */
int ht_pack_u8(const uint8_t* gt, int64_t gt_s_samp, int64_t gt_s_var, int64_t gt_s_strand,
               const uint8_t* anc, int64_t anc_s_samp, int64_t anc_s_var,
               int64_t anc_s_strand, int64_t n_samples, const ht_pack_tables* tables,
               uint32_t* planes, uint32_t* flags, int mode, void* stream);

/*
 * K1 for pgenlib chunk buffers (GenotypesPLINK._iterate, genotypes.py:1373-1402):
 * `alleles` is the int32 matrix (chunk_rows, 2*n_samples) that
 * PgenReader.read_alleles_list fills -- row r = one variant, element 2*s + c, -9 = missing
 * (never matches, like the 255 the reference turns it into at genotypes.py:1388-1390).
 * var_col[] are ROW indices into this chunk; key ids are var_key_off-relative as in
 * ht_pack_u8, offset by `key_base` so that consecutive chunks fill one plane workspace
 * of `n_keys_total` keys.
 */
int ht_pack_i32(const int32_t* alleles, int64_t row_pitch_elems, int64_t n_samples,
                const uint32_t* var_col, const uint32_t* var_key_off,
                const uint8_t* key_allele, int64_t n_vars, int64_t key_base,
                int64_t n_keys_total, uint32_t* planes, uint32_t* flags, void* stream);

/*
 * The same with local ancestry (what `transform --ancestry` does on PGEN input, haptools/transform.py:610-660,
 * without ever building gts.data or the (n, p, 2) ancestry matrix): `anc` holds the labels of THIS chunk's
 * variants, element (s, r, c) at anc + s*anc_s_samp + r*anc_s_var + c*anc_s_strand with r the chunk row
 * (var_col[]), e.g. written by ht_bp_ancestry for the chunk's positions; key_label[] as in ht_pack_tables.
 * anc == NULL (then key_label must be NULL): ht_pack_i32.
 */
int ht_pack_i32_anc(const int32_t* alleles, int64_t row_pitch_elems, int64_t n_samples, const uint8_t* anc,
                    int64_t anc_s_samp, int64_t anc_s_var, int64_t anc_s_strand, const uint32_t* var_col,
                    const uint32_t* var_key_off, const uint8_t* key_allele, const uint8_t* key_label,
                    int64_t n_vars, int64_t key_base, int64_t n_keys_total, uint32_t* planes,
                    uint32_t* flags, void* stream);

/*
 * K2 -- transform.  Replaces the per-haplotype loop
 * `hap_gts.data[:, i] = np.all(equality_arr[:, idxs[i]], axis=1)` (haplotypes.py:1411-1412;
 * with ancestry folded into the keys also transform.py:144-148).
 *
 *   hap_off[H+1], hap_key[K]  CSR of key indices per haplotype (`idxs`, haplotypes.py:1380).
 *   quad_off16[ceil(H/16)+1], quads[4*Q]  optional (both NULL or both given; `quads` 16-byte
 *           aligned): the same lists as the kernel walks them -- every list padded to whole
 *           quads of 4 by repeating its last key, entries = key * 256 (byte offset of the record
 *           inside a tile), low bits of a quad's first entry = flags (1: last quad of its
 *           haplotype, 2: haplotype without alleles), quad_off16[g] = first quad of haplotypes
 *           16 g .. 16 g + 15 (haptools_b200/plan.py quad_tables).  With them, and with `out` and
 *           `out_s_samp` multiples of 16, the half-tile kernel runs (csrc/ht_transform_half.cu:
 *           512 samples x 128 haplotypes per CTA); without them the tile kernel builds the quads
 *           from the CSR in every CTA (15-20 % slower at UKB scale).  Same result bytes either
 *           way.  Environment (measurements and tests only): HT_TRANSFORM_KERNEL=tile forces the
 *           tile kernel.  The first call on a device makes the L2 policy word the half-tile
 *           kernel takes (a one-thread kernel and a stream synchronize, once per process and
 *           device): capture CUDA graphs after one eager call.
 *   out     uint8 matrix: element (s, h, c) at out + s*out_s_samp + 2*h + c receives exactly
 *           0x00 or 0x01 (numpy bool bytes).  A haplotype without keys yields 0x01
 *           (np.all over an empty axis).  out_s_samp >= 2*H.  Rows >= n_samples are not
 *           touched.
 *   counts  optional (may be NULL) uint64[H], overwritten: number of 1-bytes of each
 *           haplotype over all samples and both strands -- the allele counts that
 *           `hap_gts.check_maf` (genotypes.py:559-625; transform.py:671-673) derives with a
 *           second pass over the result; here they fall out of the kernel (popcount).
 */
int ht_transform_u8(const uint32_t* planes, int64_t n_samples, int64_t n_keys,
                    const uint32_t* hap_off, const uint32_t* hap_key, const uint32_t* quad_off16,
                    const uint32_t* quads, int64_t n_haps, uint8_t* out, int64_t out_s_samp,
                    uint64_t* counts, void* stream);

/*
 * K2 with the key runs of the haplotype blocks: haplotypes that are neighbours in the caller's order usually
 * refer to neighbouring variants -- a .hap file that went through `haptools index` (haptools/index.py:31-94) is
 * sorted by position -- so the keys of the 128 haplotypes one CTA handles fall into a short run of consecutive
 * keys.  blk_key_lo[b] / blk_key_cnt[b] (DEVICE arrays, ceil(H / 128) entries; planner: TransformPlan.block_spans)
 * give that run for block b (cnt 0: the block has no keys).  A block whose run has at most `stage_lines`
 * (<= HT_STAGE_LINES_MAX) keys copies the run's 128-byte lines into shared memory once (cp.async) and ANDs out
 * of shared memory instead of reading one line from L2 per haplotype allele; other blocks of the same launch
 * take ht_transform_u8's path.  Needs the quads and 16-byte aligned rows (else, and with stage_lines == 0, this
 * IS ht_transform_u8).  Same result bytes.  Shared memory per CTA: 28.4 KB + stage_lines * 128 B.
 */
#define HT_STAGE_LINES_MAX 1536
int ht_transform_u8_staged(const uint32_t* planes, int64_t n_samples, int64_t n_keys,
                           const uint32_t* hap_off, const uint32_t* hap_key, const uint32_t* quad_off16,
                           const uint32_t* quads, const uint32_t* blk_key_lo, const uint32_t* blk_key_cnt,
                           int64_t stage_lines, int64_t n_haps, uint8_t* out, int64_t out_s_samp,
                           uint64_t* counts, void* stream);

/*
 * K2, bit-packed result: out_bits has the plane layout with haplotypes in place of keys
 * (out_bits[tile][hap][word][strand]); 1/8 of the bytes of the uint8 form, used for the
 * optional all-gather across sample shards and for device-side consumers.  Feeding it
 * back through ht_transform_u8 with an identity CSR unpacks it.
 */
int ht_transform_bits(const uint32_t* planes, int64_t n_samples, int64_t n_keys,
                      const uint32_t* hap_off, const uint32_t* hap_key, int64_t n_haps,
                      uint32_t* out_bits, void* stream);

/*
 * K2 bit-packed + all-gather in ONE kernel (the optional collective of the path, north_star: "NCCL is used only
 * for an optional allgather of the packed result" -- this is the NVLink-native form of that pair): every rank
 * computes the records of its own sample shard and stores each of them into the gathered result of EVERY rank,
 * dsts[0 .. n_dsts) = device addresses of the ranks' (tiles_total, n_haps, 32, 2) buffers (its own and its peers',
 * mapped through ht_peer_open), at tile dst_tile0 + local tile.  The transfer is the kernel's own stores over
 * NVLink / NVSwitch peer memory and overlaps the AND arithmetic record by record; what is left for the caller is a
 * barrier across ranks once the kernel has finished on all of them (haptools_b200.dist.PeerGather does it with a
 * stream-ordered one-element all-reduce).  n_dsts <= HT_MAX_PEERS.  With n_dsts == 1 and dst_tile0 == 0 it is
 * ht_transform_bits.  hap_order (optional, device, n_haps entries: a permutation of 0 .. n_haps-1): the order in
 * which the haplotypes are COMPUTED (the planner's key order makes neighbouring warps share plane lines); every
 * record still lands at its haplotype's own position -- 256-byte records scatter for free.
 */
#define HT_MAX_PEERS 16
int ht_transform_bits_gather(const uint32_t* planes, int64_t n_samples, int64_t n_keys,
                             const uint32_t* hap_off, const uint32_t* hap_key, const uint32_t* hap_order, int64_t n_haps,
                             uint32_t* const* dsts, int32_t n_dsts, int64_t dst_tile0, void* stream);

/*
 * Peer memory for ht_transform_bits_gather, one process per GPU: ht_peer_alloc gives device memory of the current
 * device (cudaMalloc, zero-filled) and the 64-byte handle another process of the node turns into an address of its
 * own with ht_peer_open (cudaIpcGetMemHandle / cudaIpcOpenMemHandle with lazy peer access).  ht_peer_close unmaps a
 * peer's buffer, ht_peer_free releases one's own.
 */
#define HT_PEER_HANDLE_BYTES 64
int ht_peer_alloc(size_t bytes, void** ptr, void* handle);
int ht_peer_open(const void* handle, void** ptr);
int ht_peer_close(void* ptr);
int ht_peer_free(void* ptr);

/*
 * K2, two-phase ("local") form: same results as ht_transform_u8 / ht_transform_bits with fewer
 * L2 accesses when haplotypes share alleles (K >> A).  The planner orders the haplotypes by
 * their first key and cuts them into blocks whose keys fall into a run of at most `dmax`
 * consecutive keys (haptools_b200/plan.py build_local_tables); phase 1 stages that run of
 * (tile, key) records in shared memory with one TMA bulk copy per (tile, block), ANDs every
 * haplotype of the block out of shared memory and writes the bit-packed result; phase 2
 * (ht_unpack_bits_u8) expands it to bytes.  All tables are DEVICE pointers:
 *   blk_hap_off[NB+1]  block b holds the haplotypes at sorted positions blk_hap_off[b] ..
 *                      blk_hap_off[b+1]-1 (at most 128 per block)
 *   blk_key_lo[NB], blk_key_cnt[NB]  the block's run of keys; cnt == 0: not staged (records are
 *                      read straight from the plane workspace), cnt <= dmax otherwise
 *   s_hap_off[H+1], s_hap_key[K]  CSR of the key lists in sorted order; keys of a staged block
 *                      are RELATIVE to blk_key_lo (< cnt), keys of other blocks absolute
 *   hap_order[H]       caller's index of the haplotype at each sorted position
 */
typedef struct ht_local_tables {
    int64_t n_blocks;
    int64_t dmax;
    const uint32_t* blk_hap_off;
    const uint32_t* blk_key_lo;
    const uint32_t* blk_key_cnt;
    const uint32_t* s_hap_off;
    const uint32_t* s_hap_key;
    const uint32_t* hap_order;
} ht_local_tables;

/* Phase 1 alone: the bit-packed result of ht_transform_bits (same layout, caller's haplotype
 * order), plus the optional allele counts of ht_transform_u8. */
int ht_transform_bits_local(const uint32_t* planes, int64_t n_samples, int64_t n_keys,
                            const ht_local_tables* tables, int64_t n_haps, uint32_t* out_bits,
                            uint64_t* counts, void* stream);

/* Phase 2 alone: bit-packed result (tiles, H, 32, 2) -> bytes; element (s, h, c) at
 * out + s*out_s_samp + 2*h + c receives 0x00 / 0x01.  Rows >= n_samples are not touched. */
int ht_unpack_bits_u8(const uint32_t* bits, int64_t n_samples, int64_t n_haps, uint8_t* out,
                      int64_t out_s_samp, void* stream);

/* Both phases, alternating over groups of tiles through a caller-provided workspace of
 * `workspace_bytes` (16-byte aligned, at least n_haps * 256 bytes = one tile;
 * ht_local_workspace_bytes() is the recommended size).  Arguments as ht_transform_u8. */
size_t ht_local_workspace_bytes(int64_t n_samples, int64_t n_haps);
int ht_transform_u8_local(const uint32_t* planes, int64_t n_samples, int64_t n_keys,
                          const ht_local_tables* tables, int64_t n_haps, uint8_t* out,
                          int64_t out_s_samp, uint64_t* counts, void* workspace,
                          size_t workspace_bytes, void* stream);

/*
 * Per-haplotype allele counts of a bit-packed result (popcount over samples and strands)
 * -- the numerator of GenotypesVCF.check_maf (genotypes.py:559-625) for `--maf`.
 * counts[h] (uint64) is overwritten.
 */
int ht_count_bits(const uint32_t* bits, int64_t n_samples, int64_t n_haps, uint64_t* counts,
                  void* stream);

/*
 * Hand-off to the PGEN writer: GenotypesPLINK.write (genotypes.py:1569-1659) transposes the result
 * to haplotype-major, widens a chunk of rows to int32 and flattens it to (rows, 2 n) for
 * PgenWriter.append_alleles_batch (:1618-1642).  From the bit-packed result that is a plain
 * expansion: row r of `out` (r < n_rows) is haplotype h0 + r, element 2 s + c (s < n_samples) its
 * presence on strand c of sample s, as uint8 (elem_size 1) or int32 (elem_size 4; 16-byte
 * aligned, pitch a multiple of 4).  out_pitch_elems >= 2 n_samples, in elements.  The
 * `allele_cts` argument of the writer (_num_unique_alleles, genotypes.py:1538-1567) is 2 for
 * every row of a 0/1 matrix.
 */
int ht_unpack_bits_hapmajor(const uint32_t* bits, int64_t n_samples, int64_t n_haps, int64_t h0,
                            int64_t n_rows, int elem_size, void* out, int64_t out_pitch_elems,
                            void* stream);

/*
 * Hand-off to the VCF writer: GenotypesVCF.write (genotypes.py:752-813) fills
 * record.samples[s]["GT"] sample by sample in Python and htslib prints "a|b" per sample,
 * tab-separated, newline-terminated.  For the transform's result (phased, 0/1, never missing)
 * that text is 4 bytes per sample; row r of `out` receives the n_samples * 4 bytes
 * "a|b\ta|b\t...a|b\n" of haplotype h0 + r (the nine fixed columns before it are host text).
 * out_pitch >= 4 * n_samples, in bytes.
 */
int ht_format_vcf_gt(const uint32_t* bits, int64_t n_samples, int64_t n_haps, int64_t h0,
                     int64_t n_rows, uint8_t* out, int64_t out_pitch, void* stream);

/*
 * Deterministic synthetic genotypes for the benchmark configs (not part of the
 * reference): fills `gt` (same addressing as ht_pack_u8, all three bytes of a
 * phase-column layout are NOT touched beyond c in {0,1}) from a counter-based hash of
 * (seed, sample, variant, strand); see haptools_b200/synth.py for the numpy restatement.
 *   mode 0: iid Bernoulli(1/2) alleles
 *   mode 1: "blocky": per 64-variant block each (sample, strand) copies one of 16 founder
 *           haplotypes, so multi-variant haplotypes have non-trivial frequencies
 *   missing_per_64k: expected number of 255 calls per 65536 genotype bytes.
 */
int ht_synth_gt(uint8_t* gt, int64_t s_samp, int64_t s_var, int64_t s_strand,
                int64_t n_samples, int64_t n_vars, int64_t sample_offset, uint64_t seed,
                int mode, uint32_t missing_per_64k, void* stream);

/* Synthetic local-ancestry labels (same addressing): piecewise constant along variants in
 * tracts of `tract_len` variants with a per-(sample, strand) phase, label uniform in
 * [0, n_pops). */
int ht_synth_ancestry(uint8_t* anc, int64_t s_samp, int64_t s_var, int64_t s_strand,
                      int64_t n_samples, int64_t n_vars, int64_t sample_offset, uint64_t seed,
                      uint32_t n_pops, uint32_t tract_len, void* stream);

/*
 * Re-layout of a strided genotype (or ancestry) chunk on the device:
 * dst(s, v, c) = src(s, v, c) for c in {0, 1}, with dst element (s, v, c) at
 * dst + s*d_samp + v*d_var + c.  Used by the host pipeline to turn the reference's layouts
 * that interleave the phase column (strides (3p, 3, 1) from the PGEN reader, (3, 3n, 1) from
 * the VCF reader, haptools/data/genotypes.py:183-210, 1346-1352, 557) into the dense form the
 * fast pack kernels read.
 */
int ht_relayout_u8(const uint8_t* src, int64_t s_samp, int64_t s_var, int64_t s_strand,
                   int64_t n_samples, int64_t n_vars, uint8_t* dst, int64_t d_samp, int64_t d_var,
                   void* stream);

/*
 * Local-ancestry labels from breakpoint blocks -- Breakpoints.population_array /
 * _find_blocks (haptools/data/breakpoints.py:225-330), the producer of the ancestry matrix
 * that `transform --ancestry` uses when a .bp file exists (haptools/transform.py:644-660).
 *   blk_key[B]      (chromosome index << 32) | end position of every block; the blocks of
 *                   (sample s, strand c) are blk_key[seg_off[2s+c] .. seg_off[2s+c+1]-1],
 *                   ascending.
 *   blk_pop[B]      encoded population label of every block (Breakpoints.encode).
 *   var_key[p]      (chromosome index << 32) | position of every variant.
 *   anc             output labels, element (s, v, c) at anc + s*s_samp + v*s_var + c*s_strand:
 *                   the label of the first block whose end position is >= the variant's
 *                   (np.searchsorted side="left") on the variant's chromosome, 255 if none.
 *   bad             optional uint64, initialised to ~0 by the caller: receives the minimum of
 *                   ((2s + c) << 32 | v) over the (strand, variant) pairs without a block.
 */
int ht_bp_ancestry(const uint64_t* blk_key, const uint8_t* blk_pop, const uint32_t* seg_off,
                   const uint64_t* var_key, uint8_t* anc, int64_t s_samp, int64_t s_var,
                   int64_t s_strand, int64_t n_samples, int64_t n_vars, uint64_t* bad,
                   void* stream);

/*
 * Genotype QC in one pass over the device copy of `gts.data` -- what the reference's checks derive with full
 * numpy passes on the host before transform_haps calls the transform (haptools/transform.py:623-631):
 *   Genotypes.check_missing   haptools/data/genotypes.py:444-485  (any strand >= 254; GenotypesAncestry,
 *                             haptools/transform.py:353-383: == 255 -- `missing_min`)
 *   Genotypes.check_biallelic genotypes.py:487-528  (any strand > 1)
 *   Genotypes.check_phase     genotypes.py:530-557  ((bool(g0) ^ bool(g1)) & ~bool(phase); depth == 3 only)
 *   Genotypes.check_maf       genotypes.py:559-625  (per-variant count of non-zero strand values)
 * gt element (s, v, c) at gt + s*s_samp + v*s_var + c*s_strand, c < depth (2, or 3 with the phase byte).
 *   summary           DEVICE struct; the caller initialises first_* to ~0 and the counters to 0; the kernel
 *                     atomically folds this call in, so sample chunks can be checked one after the other
 *                     (`sample_offset` = index of this chunk's first sample in the whole matrix).  first_* =
 *                     minimum of (sample << 32 | variant) over the offending pairs = the pair np.nonzero lists
 *                     first on the (n, p) mask, i.e. the one the reference's error messages name.
 *   samp_missing[n]   optional, caller-zeroed: 1 for every sample of this chunk with a missing call
 *                     (check_missing(discard_also=True) deletes exactly these)
 *   var_multiallelic[p]  optional, caller-zeroed: 1 for every variant with a value > 1 in some sample
 *   var_alt_count[p]  optional, caller-zeroed uint64: accumulated count of non-zero strand values
 */
typedef struct ht_qc_summary {
    uint64_t first_missing;
    uint64_t first_multiallelic;
    uint64_t first_unphased;
    uint64_t n_missing;       /* (sample, variant) pairs with a missing call       */
    uint64_t n_multiallelic;  /* (sample, variant) pairs with a value > 1 (check_biallelic's log line counts pairs) */
} ht_qc_summary;

int ht_qc_u8(const uint8_t* gt, int64_t s_samp, int64_t s_var, int64_t s_strand, int depth,
             int64_t n_samples, int64_t n_vars, int64_t sample_offset, int missing_min,
             ht_qc_summary* summary, uint8_t* samp_missing, uint8_t* var_multiallelic,
             uint64_t* var_alt_count, void* stream);

/*
 * Host <-> device plumbing for the Python host layer (GenotypesPLINK-style chunked input,
 * haptools/data/genotypes.py:1353-1406, and strided result read-back): a pitched 2-D copy
 * (cudaMemcpy2DAsync; `kind` = HT_COPY_*) and page-locking of an existing host range, so that
 * numpy buffers can be DMA'd without a staging copy.  h_ptr is a HOST pointer.
 * Dense buffers (both pitches == width) go out as one linear copy.  An HT_COPY_H2D source of 4 MB
 * or more in PAGEABLE host memory (what `gts.data` is when it comes from cyvcf2 / pgenlib / numpy)
 * is uploaded by the host copy engine below (submit + wait: the call returns when the data is in device
 * memory) instead of the driver's single-threaded staging; an HT_COPY_D2H destination of 4 MB or more in
 * pageable memory is filled the same way in the other direction (returns when the destination holds the data).
 */
int ht_copy2d(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch,
              int64_t width_bytes, int64_t height, int kind, void* stream);

/*
 * Upload of SELECTED rows of a host matrix: row r of `dst` receives row row_idx[r] (host array) of `src`.
 * What Genotypes.subset's column gather (genotypes.py:440) amounts to for the variant-major layout of the
 * VCF reader, where a variant is a contiguous row: only the variants some haplotype refers to cross PCIe
 * (1000G-scale config: 0.52 GB instead of 5 GB).  A PAGE-LOCKED source with 16-byte aligned rows (and dst) is
 * gathered by a kernel whose warps read the rows straight out of host memory over PCIe: asynchronous on
 * `stream`, no host thread touches a byte, the source must stay valid until the stream gets there (the
 * indices are copied during the call).  Any other source: the host copy engine gathers the rows into its
 * page-locked slots (submit + wait: the call returns when the rows are in device memory); rows must fit a
 * slot (8 MB).  HT_GATHER=bounce in the environment forces the
 * second way (measurements, tests).
 */
int ht_copy_rows_h2d(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t width_bytes,
                     const uint32_t* row_idx, int64_t n_rows, void* stream);
int ht_host_register(void* h_ptr, size_t bytes);
int ht_host_unregister(void* h_ptr);

/*
 * The result as one bit string per sample: rowbits + s * pitch holds, at bit j % 8 of byte j / 8, result byte
 * j = 2 h + c of sample s -- the bytes of the reference's `hap_gts.data[s]` (haplotypes.py:1408-1412) at one
 * bit each.  This is the form in which a result leaves the device for a HOST array: 1/8 of the bytes cross
 * PCIe, and the host only widens bits to bool bytes (ht_host_submit_download_expand).  `bits` is the packed
 * result of ht_transform_bits; pitch >= ht_rowbits_pitch(n_haps) (= 4 * ceil(H / 16) rounded up to 16), a
 * multiple of 4; bits of haplotypes >= n_haps are 0; rows >= n_samples are not touched.
 */
int64_t ht_rowbits_pitch(int64_t n_haps);
int ht_bits_to_rowbits(const uint32_t* bits, int64_t n_samples, int64_t n_haps, uint8_t* rowbits,
                       int64_t pitch, void* stream);

/*
 * Host copy engine (csrc/ht_hostcopy.cu): a persistent pool of host threads (HT_HOST_COPY_THREADS in the
 * environment, default min(cores, 12)), each with two page-locked 8 MB staging slots and its own copy
 * stream per device.  It serves the transfers that plain DMA cannot do well: PAGEABLE host memory on either
 * side (what `gts.data` and the np.empty result are in the reference's API), the gather of selected rows,
 * and the widening of the bit-packed result.  A job is asynchronous to the caller:
 *   ht_host_submit_*  records an event on `stream` (the job starts when everything queued there so far has
 *                     finished), queues the job and returns a ticket > 0 at once (-1: error, see
 *                     ht_last_error).  The host buffers must stay valid until the ticket has been waited for.
 *   ht_host_wait      returns when the job is complete: a download's destination holds the data, an
 *                     upload's data is in device memory and visible to every stream.  ticket 0 = all jobs
 *                     submitted so far.  Every ticket must be waited for once (its bookkeeping is freed then).
 * Jobs of different devices and threads share the pool; workers prefer jobs whose event has fired.
 *   upload            dst (device, pitched) <- h_src rows (h_row_idx != NULL: row r = source row h_row_idx[r];
 *                     the indices are copied during the call); rows must fit a slot (8 MB)
 *   download          h_dst (host, pitched) <- src (device)
 *   download_expand   h_out row r (n_out_bytes bytes of 0x00 / 0x01) <- bit string r of `rowbits` (device,
 *                     ht_bits_to_rowbits layout): the last step of Haplotypes.transform for a host result
 * ht_expand_rowbits_host is the widening alone, on the calling thread, for bit strings already in host memory;
 * ht_host_expand_isa(-1) tells which form runs (0 portable, 1 AVX2, 2 AVX-512BW); a value >= 0 caps it, -2
 * lifts the cap (tests).
 */
int64_t ht_host_submit_upload(void* dst, int64_t dst_pitch, const void* h_src, int64_t src_pitch,
                              int64_t width_bytes, int64_t height, const uint32_t* h_row_idx, void* stream);
int64_t ht_host_submit_download(void* h_dst, int64_t dst_pitch, const void* src, int64_t src_pitch,
                                int64_t width_bytes, int64_t height, void* stream);
int64_t ht_host_submit_download_expand(void* h_out, int64_t out_pitch, int64_t n_out_bytes,
                                       const void* rowbits, int64_t bits_pitch, int64_t n_rows, void* stream);
int ht_host_wait(int64_t ticket);
int ht_host_workers(void);
int ht_host_is_pageable(const void* h_ptr); /* 1: ordinary (unregistered) host memory, 0: page-locked / device */
int ht_expand_rowbits_host(const uint8_t* h_rowbits, int64_t bits_pitch, int64_t n_rows,
                           int64_t n_out_bytes, uint8_t* h_out, int64_t out_pitch);
int ht_host_expand_isa(int force);

#ifdef __cplusplus
}
#endif
#endif /* HAPTOOLS_CUDA_H */
