/*
 * bsalign_b200.h -- C ABI of libbsalign_b200.so: the B200 (sm_100a) implementation of bsalign's banded
 * striped DP hot path.  Plain pointers and sizes only; no CUDA, torch or reference types cross this line.
 *
 * The reference (ruanjue/bsalign) is a header-only C library: its "FFI" for this path is the pair of
 * static-inline entry points
 *     banded_striped_epi8_seqalign_pairwise   bsalign.h:3854  (declared bsalign.h:399)
 *     striped_seqedit_pairwise                bsalign.h:1046  (declared bsalign.h:232)
 * which align ONE pair per call on one CPU thread.  A GPU only pays off on batches, so the boundary is a
 * batch form of those two calls (same argument meaning, one entry per pair) plus a single-pair
 * convenience wrapper with the reference's argument list.  include/bsalign_b200_compat.h re-bodies the
 * reference's own function names on top of this ABI.
 *
 * Conventions shared with the reference:
 *   - sequences: one base per byte, values 0..3 (bsalign.h:399 `u1i *qseq`; dna.h:662 never emits 4)
 *   - mode: SEQALIGN_MODE_GLOBAL 0 / OVERLAP 1 / EXTEND 2 (bsalign.h:30-38); flag bits are ignored here
 *     (QPROF, CIGRESV are handled by the compat header on the host side)
 *   - bandwidth 0 = full band; rounded up to 16 (epi8, bsalign.h:3861-3862) / the 64-multiple rule of
 *     bsalign.h:1055-1067 (edit)
 *   - matrix[q*4+t] int8 (bsalign.h:323), gap penalties are NEGATIVE int8 (main.c:284-288)
 *   - result: the 10 ints of seqalign_result_t (bsalign.h:213-218)
 *   - cigar words: len<<4 | op, op 0=M 1=I 2=D (bsalign.h:61-69, 401-417), query-start to query-end order
 *
 * Per-pair status flags (0 = the result is bit-identical to the reference's SSE4.2 build):
 */
#ifndef BSALIGN_B200_H
#define BSALIGN_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSB200_ST_RANGE   1  /* a traceback score lookup left the stored band: the reference reads out of bounds here */
#define BSB200_ST_LOOP    2  /* no consistent traceback predecessor: the reference never terminates here (bsalign.h:3798-3814) */
#define BSB200_ST_CIGCAP  4  /* cigar capacity given by the caller was too small; cigar truncated, counts still exact */
#define BSB200_ST_REFBUG  8  /* edit band shift hits the reference's scratch-corrupting path (bsalign.h:704-713); result follows the intended algorithm */
#define BSB200_ST_EMPTY  16  /* qlen==0 or tlen==0: all-zero result (bsalign.h:1051-1054) */
#define BSB200_ST_UNSUPPORTED 32  /* edit pair with a MOVING band wider than 16384 cells (or a query of more than ~29 kb under a moving band): not built; all-zero result.
                                     Unbanded edit alignments (OVERLAP, EXTEND, GLOBAL with bandwidth 0 or >= qlen) have no length limit. */

/* same field order as seqalign_result_t, bsalign.h:213-218 */
typedef struct {
	int32_t score;
	int32_t qb, qe;
	int32_t tb, te;
	int32_t mat, mis, ins, del, aln;
} bsb200_result_t;

typedef struct bsb200_ctx bsb200_ctx;

/* kernel-time accounting of the last *_run / *_batch call, CUDA events on the context's stream */
typedef struct {
	float h2d_ms, forward_ms, traceback_ms, d2h_ms, total_ms;
	uint32_t forward_launches, traceback_launches, other_launches, waves;
	uint64_t cells;            /* nominal band cells = sum over pairs of bw_eff * tlen */
	uint64_t trace_bytes;      /* algorithmic bytes written by the forward kernel(s) */
	uint64_t h2d_bytes, d2h_bytes;
	float run_ms;              /* whole bsb200_batch_run: first memset to last kernel, one event pair */
	uint32_t reserved;
} bsb200_timing_t;

/* ---- context ------------------------------------------------------------------------------------ */
/* device: CUDA ordinal.  trace_budget_bytes: cap for the traceback store per wave (0 = 90% of free HBM). */
bsb200_ctx *bsb200_create(int device, uint64_t trace_budget_bytes);
void bsb200_destroy(bsb200_ctx *ctx);
const char *bsb200_last_error(bsb200_ctx *ctx);   /* "" when the last call succeeded */
void bsb200_trim(bsb200_ctx *ctx);                /* release the traceback arena and all cached buffers (they are re-allocated on demand) */
bsb200_ctx *bsb200_default_context(void);         /* one shared context on device 0, created on first use (used by the compat headers) */
int bsb200_device_count(void);
const char *bsb200_version(void);
void bsb200_get_timing(bsb200_ctx *ctx, bsb200_timing_t *out);

/* ---- 8-bit affine / 2-piece banded alignment: replaces banded_striped_epi8_seqalign_pairwise ------ */
/*
 * One call = n independent pairs.  seqs is one arena; pair i is query seqs[qoff[i] .. +qlen[i]) and
 * target seqs[toff[i] .. +tlen[i]).  results[n]; ncigar[n] (may be NULL); cigars is a caller arena where
 * pair i owns words [cgoff[i], cgoff[i+1]) (both may be NULL: counts only, like cigars==NULL in
 * bsalign.h:3713); status[n] may be NULL.  All pointers are HOST pointers.  Returns 0, or <0 on a CUDA /
 * argument error (message via bsb200_last_error).
 */
int bsb200_epi8_pairwise_batch(bsb200_ctx *ctx, uint64_t n, const uint8_t *seqs,
		const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		int mode, uint32_t bandwidth, const int8_t matrix[16], int8_t gapo1, int8_t gape1, int8_t gapo2, int8_t gape2,
		bsb200_result_t *results, uint32_t *cigars, const uint64_t *cgoff, uint32_t *ncigar, int32_t *status);

/* Reference argument list for ONE pair (bsalign.h:399 minus mempool/verbose); cigar_cap words at cigar. */
int bsb200_epi8_pairwise(bsb200_ctx *ctx, const uint8_t *qseq, uint32_t qlen, const uint8_t *tseq, uint32_t tlen,
		int mode, uint32_t bandwidth, const int8_t matrix[16], int8_t gapo1, int8_t gape1, int8_t gapo2, int8_t gape2,
		bsb200_result_t *result, uint32_t *cigar, uint32_t cigar_cap, uint32_t *ncigar, int32_t *status);

/* ---- 2-bit-plane edit distance: replaces striped_seqedit_pairwise --------------------------------- */
int bsb200_edit_pairwise_batch(bsb200_ctx *ctx, uint64_t n, const uint8_t *seqs,
		const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		int mode, uint32_t bandwidth,
		bsb200_result_t *results, uint32_t *cigars, const uint64_t *cgoff, uint32_t *ncigar, int32_t *status);

int bsb200_edit_pairwise(bsb200_ctx *ctx, const uint8_t *qseq, uint32_t qlen, const uint8_t *tseq, uint32_t tlen,
		int mode, uint32_t bandwidth,
		bsb200_result_t *result, uint32_t *cigar, uint32_t cigar_cap, uint32_t *ncigar, int32_t *status);

/* ---- k-mer guided edit alignment: replaces kmer_striped_seqedit_pairwise (bsalign.h:1209-1536; main.c:196 `bsalign edit -m kmer -k ksz`,
 * bspoa.h:2089) ----
 * ksz: k-mer size, 1..15 (larger values are clamped to 15 like bsalign.h:1217).  Unique shared canonical k-mers anchor the pair, the
 * gaps between the anchors are aligned with the edit DP, a pair without usable anchors gets the plain global edit (bsalign.h:1440).
 * Arguments as bsb200_edit_pairwise_batch; a pair needs room for qlen + tlen + 2 cigar words.  The *_dense form returns the cigars
 * dense and in pair order like bsb200_pairwise_batch_dense. */
int bsb200_kmer_edit_batch(bsb200_ctx *ctx, uint64_t n, const uint8_t *seqs,
		const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen, uint32_t ksz,
		bsb200_result_t *results, uint32_t *cigars, const uint64_t *cgoff, uint32_t *ncigar, int32_t *status);
int bsb200_kmer_edit_batch_dense(bsb200_ctx *ctx, uint64_t n, const uint8_t *seqs,
		const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen, uint32_t ksz,
		bsb200_result_t *results, uint32_t *cigars, uint64_t cigar_cap_words, uint64_t *total_words, uint32_t *ncigar, int32_t *status);
/* the reference's argument list for ONE pair (bsalign.h:1209 minus mempool / verbose) */
int bsb200_kmer_edit_pairwise(bsb200_ctx *ctx, uint32_t ksz, const uint8_t *qseq, uint32_t qlen, const uint8_t *tseq, uint32_t tlen,
		bsb200_result_t *result, uint32_t *cigar, uint32_t cigar_cap, uint32_t *ncigar, int32_t *status);

/* ---- re-alignment of reads against the MSA profile: the DP and the walk of remsa_pedit_rd_bspoacore (bspoa.h:3916-4045, kernel
 * maxmat_dp_diag_rowcal bspoa.h:3856) for batches of (object, read) jobs ----
 * hdr: per job 8 ints {mlen, bw, mbeg, mend, rend, 0, 0, 0} (bw = the band in cells, a multiple of 16, <= 256; rend = the read positions).
 * `in` holds, at in_off[job], the job's ten arrays exactly as remsa_pedits_bspoa lays them out in g->memp (bspoa.h:4209-4229): seqs[0],
 * seqs[1], mats[0][0..3], mats[1][0..3], each roundup(mlen + bw, 16) bytes with bw / 2 bytes of padding in front of index 0.
 * match: rend ints per job at match_off[job] - the MSA column every read position is matched to (the reference calls merge_nodes_bspoa
 * for exactly those pairs, bspoa.h:4012-4023), -1 otherwise.  out: 4 ints per job {score of the walk, status (0 = fine), matched
 * positions, 0}.  matrices / mat_off (may be NULL): the two DP matrices per job, (2 * mlen + 1) * (bw + 2) bytes each, for checks. */
int bsb200_remsa_batch(bsb200_ctx *ctx, uint32_t njobs, const int32_t *hdr, const uint8_t *in, const uint64_t *in_off, uint64_t in_bytes,
		int32_t *match, const uint64_t *match_off, uint64_t match_ints, int32_t *out, uint8_t *matrices, const uint64_t *mat_off);

/* ---- more shapes of the same call (kind: 0 = epi8, 1 = edit; matrix / gaps ignored for kind 1) ------------------------------------ */
/* one call, cigars DENSE and in pair order (pair i starts at word sum(ncigar[0..i-1]); see bsb200_batch_fetch_dense): the fast path */
int bsb200_pairwise_batch_dense(bsb200_ctx *ctx, int kind, uint64_t n, const uint8_t *seqs,
		const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		int mode, uint32_t bandwidth, const int8_t matrix[16], int8_t gapo1, int8_t gape1, int8_t gapo2, int8_t gape2,
		bsb200_result_t *results, uint32_t *cigars, uint64_t cigar_cap_words, uint64_t *total_words, uint32_t *ncigar, int32_t *status);
/* the same call with the sequences 2-BIT PACKED (a BaseBank's words as in bsb200_batch_upload_bits; qoff / toff are base offsets).  Both
 * forms pipeline large edit batches internally: chunks of pairs go through two streams, so copies, host planning and kernels overlap. */
int bsb200_pairwise_batch_dense_bits(bsb200_ctx *ctx, int kind, uint64_t n, const uint64_t *bits,
		const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		int mode, uint32_t bandwidth, const int8_t matrix[16], int8_t gapo1, int8_t gape1, int8_t gapo2, int8_t gape2,
		bsb200_result_t *results, uint32_t *cigars, uint64_t cigar_cap_words, uint64_t *total_words, uint32_t *ncigar, int32_t *status);
/* one POINTER per sequence, the way the reference's callers hold them (bsalign.h:399 takes u1i *qseq, u1i *tseq per call); gathered into
 * one pinned arena by `nthreads` host threads.  cigar_out (may be NULL) holds one pointer per pair (entries may be NULL) with room for
 * qlen[i] + tlen[i] + 2 words. */
int bsb200_pairwise_batch_ptrs(bsb200_ctx *ctx, int kind, uint64_t n, const uint8_t *const *q, const uint32_t *qlen,
		const uint8_t *const *t, const uint32_t *tlen, int mode, uint32_t bandwidth, const int8_t matrix[16],
		int8_t gapo1, int8_t gape1, int8_t gapo2, int8_t gape2, bsb200_result_t *results, uint32_t *const *cigar_out, uint32_t *ncigar, int32_t *status, int nthreads);
/* every GPU of the box from ONE process: ctxs[d] = bsb200_create(d, 0).  The batch is cut into shards of equal DP cells, one host thread
 * per device packs and runs its shard; same arguments and results as bsb200_epi8_pairwise_batch / bsb200_edit_pairwise_batch. */
int bsb200_pairwise_batch_multi(bsb200_ctx *const *ctxs, int nctx, int kind, uint64_t n, const uint8_t *seqs,
		const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		int mode, uint32_t bandwidth, const int8_t matrix[16], int8_t gapo1, int8_t gape1, int8_t gapo2, int8_t gape2,
		bsb200_result_t *results, uint32_t *cigars, const uint64_t *cgoff, uint32_t *ncigar, int32_t *status);

/* ---- staged form (inputs resident in HBM): upload once, run the kernels many times, fetch ---------- */
typedef struct bsb200_batch bsb200_batch;
/* kind: 0 = epi8, 1 = edit.  matrix/gaps are ignored for kind 1. */
bsb200_batch *bsb200_batch_upload(bsb200_ctx *ctx, int kind, uint64_t n, const uint8_t *seqs,
		const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		int mode, uint32_t bandwidth, const int8_t matrix[16], int8_t gapo1, int8_t gape1, int8_t gapo2, int8_t gape2,
		int want_cigar);
/* Same, with the sequences 2-BIT PACKED as the reference stores them: `bits` is a BaseBank's word array (dna.h:63 bits2bit: base i =
 * bits[i >> 5] >> (((~i) & 31) << 1) & 3; main.c keeps every read that way, seqs->rdseqs, and unpacks a pair with bitseq_basebank,
 * dna.h:769, before each call); qoff / toff are BASE offsets into it (seqs->rdoffs).  A quarter of the bytes cross PCIe; the bases
 * are unpacked on the device. */
bsb200_batch *bsb200_batch_upload_bits(bsb200_ctx *ctx, int kind, uint64_t n, const uint64_t *bits,
		const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		int mode, uint32_t bandwidth, const int8_t matrix[16], int8_t gapo1, int8_t gape1, int8_t gapo2, int8_t gape2,
		int want_cigar);
int bsb200_batch_run(bsb200_ctx *ctx, bsb200_batch *b);       /* kernels only; returns when they are done */
int bsb200_batch_sync(bsb200_ctx *ctx);                       /* wait for the stream */
int bsb200_batch_fetch(bsb200_ctx *ctx, bsb200_batch *b, bsb200_result_t *results,
		uint32_t *cigars, const uint64_t *cgoff, uint32_t *ncigar, int32_t *status);
/* Same, but the cigars come back DENSE and in pair order: pair i starts at word sum(ncigar[0..i-1]) of `cigars`
 * (capacity cigar_cap_words; *total_words receives the number of words written, or needed when the call fails).
 * This is the fast path: the ordering is done on the device and lands in the caller's buffer with one copy. */
int bsb200_batch_fetch_dense(bsb200_ctx *ctx, bsb200_batch *b, bsb200_result_t *results,
		uint32_t *cigars, uint64_t cigar_cap_words, uint64_t *total_words, uint32_t *ncigar, int32_t *status);
void bsb200_batch_free(bsb200_ctx *ctx, bsb200_batch *b);

/* ---- multi-GPU split (one process per GPU; the batch lives on rank 0) ------------------------------------------
 * The reference has no notion of devices: banded_striped_epi8_seqalign_pairwise (bsalign.h:3854) is called pair by pair, and pairs
 * never interact.  A batch is therefore cut into shards of equal DP cells, each shard's compact arena travels from rank 0's GPU to
 * its rank's GPU (NCCL over NVLink, bsalign_b200/shard.py), is aligned there in place and the results travel back the same way.
 * These entry points are that path's ends on a rank: the sequence arena / the results are DEVICE memory of the caller, the small
 * offset and length tables stay host arrays (every rank derives them from the broadcast pair lengths). */
bsb200_batch *bsb200_batch_upload_dev(bsb200_ctx *ctx, int kind, uint64_t n, const uint8_t *d_seqs /* device; kept alive by the caller until batch_free */,
		const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		int mode, uint32_t bandwidth, const int8_t matrix[16], int8_t gapo1, int8_t gape1, int8_t gapo2, int8_t gape2,
		int want_cigar);
/* d_results n x 10 int32, d_ncigar n, d_status n (kernel flags only; BSB200_ST_EMPTY is added by the host-side fetches), d_cigars dense
 * and pair-ordered like bsb200_batch_fetch_dense; all DEVICE pointers, any may be NULL.  *total_words always receives the word count. */
int bsb200_batch_fetch_dense_dev(bsb200_ctx *ctx, bsb200_batch *b, int32_t *d_results, uint32_t *d_cigars, uint64_t cigar_cap_words,
		uint64_t *total_words, uint32_t *d_ncigar, int32_t *d_status);
/* host helpers of the split: pairs idx[0..m) copied into one compact arena (query, then target, in idx order; offsets to out_qoff /
 * out_toff; out_seqs NULL: only size it), and the pair-ordered merge of the shards' dense cigars */
uint64_t bsb200_pack_pairs(const uint8_t *seqs, const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		const uint64_t *idx, uint64_t m, uint8_t *out_seqs, uint64_t *out_qoff, uint64_t *out_toff, int nthreads);
/* the same packing on the device: d_src is the caller's whole arena in this device's memory, d_dst receives the compact arena
 * (bsb200_pack_pairs with out_seqs == NULL sizes it); all other pointers are host arrays */
int bsb200_pack_pairs_dev(bsb200_ctx *ctx, const uint8_t *d_src, const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		const uint64_t *idx, uint64_t m, uint8_t *d_dst);
void bsb200_scatter_words(uint32_t *dst, const uint64_t *dst_off, const uint32_t *src, const uint64_t *src_off, const uint32_t *len, uint64_t m, int nthreads);

/* ---- ingest / egress next to the path (host code): sequence files into BaseBank words, and the command line's text ------------------ */
/* readseq_filereader (filereader.h:609) + seq2basebank (dna.h:653-671): FASTA / FASTQ, plain or .gz; tag = header up to the first blank;
 * every base becomes base_bit_table[c] & 3 (non-ACGT -> A), first base of a word in its top two bits; empty records are dropped like
 * main.c:312 does.  Check bsb200_seqfile_error ("" = fine). */
typedef struct bsb200_seqfile bsb200_seqfile;
bsb200_seqfile *bsb200_seqfile_read(const char *path);
const char *bsb200_seqfile_error(const bsb200_seqfile *sf);
uint64_t bsb200_seqfile_nseq(const bsb200_seqfile *sf);
uint64_t bsb200_seqfile_nbases(const bsb200_seqfile *sf);
const uint64_t *bsb200_seqfile_bits(const bsb200_seqfile *sf);       /* BaseBank words (+ one spare word) */
const uint64_t *bsb200_seqfile_offsets(const bsb200_seqfile *sf);    /* base offset of every record */
const uint32_t *bsb200_seqfile_lengths(const bsb200_seqfile *sf);
const char *bsb200_seqfile_name(const bsb200_seqfile *sf, uint64_t i);
void bsb200_seqfile_free(bsb200_seqfile *sf);
/* seqalign_cigar2alnstr (bsalign.h:531-582) on BaseBank words: query row, target row, match row ('|' '*' '-'); each needs length + 1 bytes */
uint32_t bsb200_cigar2alnstr(const uint64_t *bits, uint64_t qoff, uint64_t toff, const bsb200_result_t *rs, const uint32_t *cigar, uint32_t ncigar,
		char *qrow, char *trow, char *mrow, uint32_t length);
/* the text `bsalign align` / `bsalign edit` print for one pair (main.c:346-347 / 226-228 + the three rows; nothing when rs->mat == 0);
 * returns the bytes needed, writes them when they fit cap */
uint64_t bsb200_format_pair_text(char *out, uint64_t cap, const char *qname, uint32_t qlen, const char *tname, uint32_t tlen, const bsb200_result_t *rs,
		const uint64_t *bits, uint64_t qoff, uint64_t toff, const uint32_t *cigar, uint32_t ncigar);
/* the whole command: consecutive records of `path` are pairs (main.c:314), aligned in batches through bsb200_batch_upload_bits; the
 * reference's text goes to `out` (a FILE*) in input order.  kind 0 = align, 1 = edit, 2 = `edit -m kmer` (bandwidth carries the k-mer
 * size).  Returns the number of pairs, or -1. */
int64_t bsb200_align_file(bsb200_ctx *ctx, int kind, const char *path, int mode, uint32_t bandwidth, const int8_t matrix[16],
		int8_t gapo1, int8_t gape1, int8_t gapo2, int8_t gape2, void *out /* FILE* */, uint64_t batch_pairs /* 0 = 1M */);

/* binary MSA of the reference (dump_binary_msa_bspoa / load_binary_msa_bspoa_core, bspoa.h:1555-1643): [0x81 u32 len, metadata] 0x22 u32 mlen
 * u32 nseq, mlen columns of nseq + 1 bytes (read bases 0..3, 4 = gap; then the consensus base), mlen quality bytes, mlen alternative-base bytes,
 * 0xFF.  `out` / `inp` are FILE*; bsb200_msa_read returns the next MSA of the stream or NULL.  include/bsalign_b200_poa_compat.h wraps them
 * for a BSPOA (b200_dump_binary_msa_bspoa). */
typedef struct bsb200_msa bsb200_msa;
int bsb200_msa_write(void *out, uint32_t nseq, uint32_t mlen, const uint8_t *cols, const uint8_t *qlt, const uint8_t *alt, const char *meta, uint32_t metalen);
bsb200_msa *bsb200_msa_read(void *inp);
uint32_t bsb200_msa_nseq(const bsb200_msa *m);
uint32_t bsb200_msa_mlen(const bsb200_msa *m);
const uint8_t *bsb200_msa_cols(const bsb200_msa *m);
const uint8_t *bsb200_msa_qlt(const bsb200_msa *m);
const uint8_t *bsb200_msa_alt(const bsb200_msa *m);
const char *bsb200_msa_meta(const bsb200_msa *m, uint32_t *len);
void bsb200_msa_free(bsb200_msa *m);

/* ---- POA read-vs-graph banded DP sweep: replaces align_rd_bspoacore (bspoa.h:2515-2618) -------------------- */
/*
 * One SWEEP JOB = one call of the reference's align_rd_bspoacore: a read (query) against the selected sub-graph
 * of a BSPOA, i.e. everything between prepare_rd_align_bspoa (bspoa.h:2020-2230) and alignment2graph_bspoa
 * (bspoa.h:2274).  It covers dpalign_row_update_bspoa (bspoa.h:2232-2261: row_movx + piecex_row_cal),
 * dpalign_row_merge_bspoa (bspoa.h:2263-2272: piecex_row_merge, bsalign.h:2474-2616), the head row_init
 * (bspoa.h:2224-2226) and the end-candidate rules (bspoa.h:2548-2602).  A batch holds many independent jobs
 * (many BSPOA objects stepped in lock-step by the host).
 *
 * Job i:
 *   par[10*i ..]   = { g->bandwidth (multiple of 16), par->alnmode, M, X, O, E, Q, P, T, refbonus }   (bspoa.h:55-76)
 *   query          = queries[qoff[i] .. +slen[i])       g->qseq->buffer + g->qb, g->slen bases 0..3
 *   nodes          = [node_off[i], node_off[i+1])        LOCAL id n = position in g->sels; per node:
 *                    node_base (bspoanode_t.base), node_bonus (.bonus), node_rpos (.rpos, band offset), node_nct (.nct)
 *   out-edges      = CSR: node n owns edst[edge_off[i] + eoff[node_off[i] + i + n] .. eoff[node_off[i] + i + n + 1])
 *                    (eoff has nnode+1 entries per job), destinations are LOCAL ids, listed in the order of the
 *                    node's edge list (bspoanode_t.edge / bspoaedge_t.next) and restricted to selected nodes
 *                    (get_bitvec(g->states, e->node), bspoa.h:2538)
 *   head[i], tail[i] LOCAL ids of nhead / ntail
 * Outputs:
 *   rows: for node n of job i, bsb200_poa_block_bytes(par_i) bytes at row_off[i] + n * block_bytes, laid out exactly
 *         like the reference's g->memp block of that node (dpalign_row_prepare_data, bspoa.h:1787-1793:
 *         [u bw][e bw][q bw][ubegs 17 x int32], striped byte order), so that the arena of job i can be copied to
 *         g->memp->buffer + 2 * g->mmblk in one memcpy.  Blocks of nodes the sweep never reached are left untouched.
 *   best[3*i ..]  = { g->maxscr, g->maxidx as LOCAL id (-1: none), g->maxoff }
 *   ops[2*i ..]   = { row updates, row merges } executed (the GCUPS denominator is updates * bandwidth)
 *   status[i]     = BSB200_ST_RANGE if an end-candidate lookup left the band (the reference reads out of bounds)
 */
typedef struct bsb200_poa_batch bsb200_poa_batch;
uint32_t bsb200_poa_block_bytes(const int32_t par[10]);    /* g->mmblk, bspoa.h:2217 */
bsb200_poa_batch *bsb200_poa_upload(bsb200_ctx *ctx, uint32_t njobs, const int32_t *par,
		const uint8_t *queries, const uint64_t *qoff, const uint32_t *slen,
		const uint64_t *node_off, const uint8_t *node_base, const uint8_t *node_bonus, const int32_t *node_rpos, const int32_t *node_nct,
		const int32_t *eoff, const uint64_t *edge_off, const int32_t *edst, const uint32_t *head, const uint32_t *tail);
int bsb200_poa_run(bsb200_ctx *ctx, bsb200_poa_batch *b);
uint64_t bsb200_poa_rows_bytes(bsb200_poa_batch *b, uint64_t *row_off_out /* njobs + 1, may be NULL */);
int bsb200_poa_fetch(bsb200_ctx *ctx, bsb200_poa_batch *b, uint8_t *rows, int32_t *best, int32_t *status, uint64_t *ops); /* any may be NULL */
void bsb200_poa_free(bsb200_ctx *ctx, bsb200_poa_batch *b);
/* one-shot: upload + run + fetch with HOST buffers */
int bsb200_poa_rows_batch(bsb200_ctx *ctx, uint32_t njobs, const int32_t *par,
		const uint8_t *queries, const uint64_t *qoff, const uint32_t *slen,
		const uint64_t *node_off, const uint8_t *node_base, const uint8_t *node_bonus, const int32_t *node_rpos, const int32_t *node_nct,
		const int32_t *eoff, const uint64_t *edge_off, const int32_t *edst, const uint32_t *head, const uint32_t *tail,
		uint8_t *rows, int32_t *best, int32_t *status, uint64_t *ops);

/* ---- POA: the walk of alignment2graph_bspoa (bspoa.h:2274-2497) on the device, right behind the sweep ------------------- */
/*
 * The reference walks back from (maxidx, maxoff) to the head, re-deriving every step from the node rows, and interleaves graph
 * surgery (merge_nodes_bspoa) that never changes what the walk reads.  With the reverse edges of the selected sub-graph attached,
 * the run also performs that walk and returns its DECISIONS instead of the row blocks (15 GB per 1000-job round):
 *   reverse edges of job i: node n owns entries [reoff[node_off[i] + i + n], reoff[.. + n + 1]) (+ redge_off[i]) of resrc / recov:
 *         the nodes of its erev list (bspoanode_t.erev / bspoaedge_t.next order, restricted to selected nodes) as LOCAL ids, and
 *         bspoaedge_t.cov of every entry (the tie-break of bspoa.h:2457-2464)
 *   match: int32 per read position, job i at match[qoff[i] .. + slen[i]): LOCAL id of the node the position is aligned to
 *         (the `u->cpos = n->cpos` / merge_nodes_bspoa step of bspoa.h:2393-2406) or -1 (insertion / outside the alignment)
 *   trace[8*i ..] = { x at the end of the walk (rs.qb before + g->qb), node at the end, mat, mis, ins, del, start node, flags }
 *         flags: BSB200_ST_RANGE / BSB200_ST_LOOP as for the pairwise traceback (the reference reads out of bounds / never ends)
 * include/bsalign_b200_poa_compat.h replays merge_nodes_bspoa and the cpos bookkeeping from `match` in the reference's order.
 */
int bsb200_poa_attach_reverse(bsb200_ctx *ctx, bsb200_poa_batch *b, const int32_t *reoff, const uint64_t *redge_off,
		const int32_t *resrc, const int32_t *recov);      /* after bsb200_poa_upload, before bsb200_poa_run */
int bsb200_poa_fetch_trace(bsb200_ctx *ctx, bsb200_poa_batch *b, int32_t *match, int32_t *trace);
int bsb200_poa_align_batch(bsb200_ctx *ctx, uint32_t njobs, const int32_t *par,
		const uint8_t *queries, const uint64_t *qoff, const uint32_t *slen,
		const uint64_t *node_off, const uint8_t *node_base, const uint8_t *node_bonus, const int32_t *node_rpos, const int32_t *node_nct,
		const int32_t *eoff, const uint64_t *edge_off, const int32_t *edst, const uint32_t *head, const uint32_t *tail,
		const int32_t *reoff, const uint64_t *redge_off, const int32_t *resrc, const int32_t *recov,
		uint8_t *rows /* may be NULL */, int32_t *best, int32_t *status, uint64_t *ops, int32_t *match, int32_t *trace);

/* nominal band width the kernels use for one pair (the GCUPS denominator) */
uint32_t bsb200_epi8_bandwidth(uint32_t qlen, uint32_t bandwidth);
uint32_t bsb200_edit_bandwidth(uint32_t qlen, uint32_t tlen, int mode, uint32_t bandwidth);

#ifdef __cplusplus
}
#endif
#endif
