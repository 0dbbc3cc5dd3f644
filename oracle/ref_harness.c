/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin C entry points around the UNMODIFIED reference headers, which are compiled from where they
 * lie under /root/reference (see oracle/Makefile: -I$(REF)); no reference source is copied here.
 * The resulting oracle/_ref/libbsref.so is used
 *   (1) to pin oracle/bsalign_oracle.c (our own CPU restatement) and
 *   (2) as the "reference" CPU arm of bench.py (--impl reference, cpu_baseline.kind="reference").
 *
 * Functions wrapped:
 *   banded_striped_epi8_seqalign_pairwise   bsalign.h:3854
 *   striped_seqedit_pairwise                bsalign.h:1046
 *   kmer_striped_seqedit_pairwise           bsalign.h:1209
 * The reference is single-threaded; the *_batch entry points run a pthread pool over independent
 * pairs, one b1v mempool + u4v cigars per thread.
 */
#include "bsalign.h"
#include <pthread.h>
#include <stdint.h>

typedef struct {
	int kind; // 0 = epi8, 1 = edit
	uint64_t n;
	const uint8_t *seqs;
	const uint64_t *qoff, *toff;
	const uint32_t *qlen, *tlen;
	int mode;
	uint32_t bandwidth;
	const int8_t *matrix;
	int8_t go1, ge1, go2, ge2;
	int32_t *results;       // n * 10
	uint32_t *cigars;       // arena, may be NULL
	const uint64_t *cgoff;  // n + 1 offsets into arena (capacity per pair), may be NULL
	uint32_t *ncigar;       // n, may be NULL
	volatile uint64_t next;
	int repeat;
} bsref_job_t;

static void bsref_store_result(int32_t *out, seqalign_result_t *rs){
	out[0] = rs->score; out[1] = rs->qb; out[2] = rs->qe; out[3] = rs->tb; out[4] = rs->te;
	out[5] = rs->mat; out[6] = rs->mis; out[7] = rs->ins; out[8] = rs->del; out[9] = rs->aln;
}

static void* bsref_worker(void *arg){
	bsref_job_t *job = (bsref_job_t*)arg;
	b1v *mempool = adv_init_b1v(1024, 0, WORDSIZE, 0);
	u4v *cigars = init_u4v(64);
	seqalign_result_t rs;
	b1i mtx[16];
	uint64_t i, j, cap;
	int r;
	if(job->matrix) memcpy(mtx, job->matrix, 16);
	while(1){
		i = __sync_fetch_and_add(&job->next, 16);
		if(i >= job->n) break;
		for(j=i;j<i+16&&j<job->n;j++){
			for(r=0;r<job->repeat;r++){
				if(job->kind == 0){
					rs = banded_striped_epi8_seqalign_pairwise((u1i*)job->seqs + job->qoff[j], job->qlen[j], (u1i*)job->seqs + job->toff[j], job->tlen[j],
						mempool, cigars, job->mode, job->bandwidth, mtx, job->go1, job->ge1, job->go2, job->ge2, 0);
				} else if(job->kind == 1){
					rs = striped_seqedit_pairwise((u1i*)job->seqs + job->qoff[j], job->qlen[j], (u1i*)job->seqs + job->toff[j], job->tlen[j],
						job->mode, job->bandwidth, mempool, cigars, 0);
				} else { // k-mer guided edit (bsalign.h:1209); it reverses a prefix of both sequences in place and restores it
					rs = kmer_striped_seqedit_pairwise((u1i)job->bandwidth, (u1i*)job->seqs + job->qoff[j], job->qlen[j], (u1i*)job->seqs + job->toff[j], job->tlen[j],
						mempool, cigars, 0);
				}
			}
			bsref_store_result(job->results + j * 10, &rs);
			if(job->ncigar) job->ncigar[j] = cigars->size;
			if(job->cigars && job->cgoff){
				cap = job->cgoff[j + 1] - job->cgoff[j];
				if(cap > cigars->size) cap = cigars->size;
				memcpy(job->cigars + job->cgoff[j], cigars->buffer, cap * sizeof(uint32_t));
			}
		}
	}
	free_b1v(mempool);
	free_u4v(cigars);
	return NULL;
}

static int bsref_run(bsref_job_t *job, int nthreads){
	pthread_t *tids;
	int i;
	if(nthreads < 1) nthreads = 1;
	job->next = 0;
	if(job->repeat < 1) job->repeat = 1;
	if(nthreads == 1){
		bsref_worker(job);
		return 0;
	}
	tids = malloc(sizeof(pthread_t) * nthreads);
	for(i=0;i<nthreads;i++) pthread_create(tids + i, NULL, bsref_worker, job);
	for(i=0;i<nthreads;i++) pthread_join(tids[i], NULL);
	free(tids);
	return 0;
}

int bsref_wordsize(void){ return WORDSIZE; }

int bsref_epi8_batch(uint64_t n, const uint8_t *seqs, const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		int mode, uint32_t bandwidth, const int8_t matrix[16], int8_t go1, int8_t ge1, int8_t go2, int8_t ge2,
		int32_t *results, uint32_t *cigars, const uint64_t *cgoff, uint32_t *ncigar, int nthreads, int repeat){
	bsref_job_t job;
	memset(&job, 0, sizeof(job));
	job.kind = 0; job.n = n; job.seqs = seqs; job.qoff = qoff; job.qlen = qlen; job.toff = toff; job.tlen = tlen;
	job.mode = mode; job.bandwidth = bandwidth; job.matrix = matrix; job.go1 = go1; job.ge1 = ge1; job.go2 = go2; job.ge2 = ge2;
	job.results = results; job.cigars = cigars; job.cgoff = cgoff; job.ncigar = ncigar; job.repeat = repeat;
	return bsref_run(&job, nthreads);
}

int bsref_edit_batch(uint64_t n, const uint8_t *seqs, const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		int mode, uint32_t bandwidth, int32_t *results, uint32_t *cigars, const uint64_t *cgoff, uint32_t *ncigar, int nthreads, int repeat){
	bsref_job_t job;
	memset(&job, 0, sizeof(job));
	job.kind = 1; job.n = n; job.seqs = seqs; job.qoff = qoff; job.qlen = qlen; job.toff = toff; job.tlen = tlen;
	job.mode = mode; job.bandwidth = bandwidth;
	job.results = results; job.cigars = cigars; job.cgoff = cgoff; job.ncigar = ncigar; job.repeat = repeat;
	return bsref_run(&job, nthreads);
}

/* kmer_striped_seqedit_pairwise bsalign.h:1209 (main.c:196 `bsalign edit -m kmer -k ksz`); pairs must not share bytes (in-place reversal) */
int bsref_kmer_edit_batch(uint64_t n, uint8_t *seqs, const uint64_t *qoff, const uint32_t *qlen, const uint64_t *toff, const uint32_t *tlen,
		uint32_t ksz, int32_t *results, uint32_t *cigars, const uint64_t *cgoff, uint32_t *ncigar, int nthreads, int repeat){
	bsref_job_t job;
	memset(&job, 0, sizeof(job));
	job.kind = 2; job.n = n; job.seqs = seqs; job.qoff = qoff; job.qlen = qlen; job.toff = toff; job.tlen = tlen;
	job.mode = 0; job.bandwidth = ksz;
	job.results = results; job.cigars = cigars; job.cgoff = cgoff; job.ncigar = ncigar; job.repeat = repeat;
	return bsref_run(&job, nthreads);
}

/*
 * Debug aid for pinning the forward pass: run one epi8 alignment and copy out, per target row,
 * the band offset (begs), the 17 ubegs ints and the u/e/q bytes in LINEAR band order (de-striped).
 * The carving below mirrors bsalign.h:3875-3913 only to locate the rows inside the caller's pool.
 */
int bsref_epi8_rows(const uint8_t *q, uint32_t qlen, const uint8_t *t, uint32_t tlen, int mode, uint32_t bandwidth,
		const int8_t matrix[16], int8_t go1, int8_t ge1, int8_t go2, int8_t ge2,
		int32_t *result, int32_t *begs_out, int32_t *ubegs_out, int8_t *u_out, int8_t *e_out, int8_t *q_out){
	b1v *mempool = adv_init_b1v(1024, 0, WORDSIZE, 0);
	u4v *cigars = init_u4v(64);
	seqalign_result_t rs;
	b1i mtx[16], *memp, *ups, *eps, *qps, *ubs;
	u4i bw, W, i, p;
	int piecewise, *begs;
	memcpy(mtx, matrix, 16);
	rs = banded_striped_epi8_seqalign_pairwise((u1i*)q, qlen, (u1i*)t, tlen, mempool, cigars, mode, bandwidth, mtx, go1, ge1, go2, ge2, 0);
	bsref_store_result(result, &rs);
	bw = bandwidth? bandwidth : qlen;
	bw = roundup_times(bw, WORDSIZE);
	W = bw / WORDSIZE;
	piecewise = banded_striped_epi8_seqalign_get_piecewise(go1, ge1, go2, ge2, bw);
	memp = mempool->buffer + mempool->size + WORDSIZE;
	memp += banded_striped_epi8_seqalign_qprof_size(qlen, bw);
	ups = memp + bw; memp += bw * Int64(tlen + 1);
	eps = NULL; qps = NULL;
	if(piecewise){ eps = memp + bw; memp += bw * Int64(tlen + 1); }
	if(piecewise == 2){ qps = memp + bw; memp += bw * Int64(tlen + 1); }
	ubs = memp + roundup_times((WORDSIZE + 1) * sizeof(int), WORDSIZE); memp += (tlen + 1) * roundup_times((WORDSIZE + 1) * sizeof(int), WORDSIZE);
	memp += roundup_times((WORDSIZE + 1) * sizeof(int), WORDSIZE);
	memp += bw * (piecewise + 1);
	begs = ((int*)memp) + 1;
	for(i=0;i<tlen;i++){
		begs_out[i] = begs[i];
		memcpy(ubegs_out + i * 17, ubs + i * roundup_times((WORDSIZE + 1) * sizeof(int), WORDSIZE), 17 * sizeof(int));
		for(p=0;p<bw;p++){
			u4i idx = (p % W) * WORDSIZE + (p / W);
			u_out[(size_t)i * bw + p] = ups[(size_t)i * bw + idx];
			if(e_out) e_out[(size_t)i * bw + p] = eps? eps[(size_t)i * bw + idx] : 0;
			if(q_out) q_out[(size_t)i * bw + p] = qps? qps[(size_t)i * bw + idx] : 0;
		}
	}
	free_b1v(mempool);
	free_u4v(cigars);
	return piecewise;
}

/* ------------------------------------------------------------------------------------------------------------
 * POA read-vs-graph sweep (align_rd_bspoacore, bspoa.h:2515-2618): run a whole POA job through the reference and
 * dump, for every read alignment, what the sweep consumes (query, selected sub-graph in the reference's own edge
 * order, per-node band offsets/base/bonus/in-degree, parameters) and what it produces (every node row, the best end).
 * The loop below only CALLS the reference's functions in the order end_bspoa (bspoa.h:4722-4778) and align_rd_bspoa
 * (bspoa.h:2620-2667) call them; the dump sits between prepare_rd_align_bspoa and align_rd_bspoacore.
 * ------------------------------------------------------------------------------------------------------------ */
#include "bspoa.h"

typedef struct { uint8_t *buf; size_t n, cap; } bsref_blob_t;
static void blob_put(bsref_blob_t *b, const void *p, size_t len){
	size_t pad = (4 - (len & 3)) & 3;
	if(b->n + len + pad > b->cap){ b->cap = (b->n + len + pad) * 2 + 4096; b->buf = realloc(b->buf, b->cap); }
	memcpy(b->buf + b->n, p, len); b->n += len;
	if(pad){ memset(b->buf + b->n, 0, pad); b->n += pad; }
}

static void bsref_poa_dump_job(BSPOA *g, BSPOAPar *par, u4i nhead, u4i ntail, bsref_blob_t *blob, int phase, size_t *hdr_at){
	u4i i, nloc = g->sels->size, bw = g->bandwidth, W = bw / WORDSIZE, p;
	static u4i *loc = NULL; static size_t loccap = 0;
	if(loccap < g->nodes->size){ loccap = g->nodes->size * 2 + 16; loc = realloc(loc, loccap * sizeof(u4i)); }
	for(i=0;i<nloc;i++) loc[g->sels->buffer[i]] = i;
	if(phase == 0){
		int32_t hdr[24]; u4i nedge = 0, eidx;
		int32_t *node = malloc(sizeof(int32_t) * 5 * nloc), *eoff = malloc(sizeof(int32_t) * (nloc + 1)), *edst;
		for(i=0;i<nloc;i++){
			bspoanode_t *u = ref_bspoanodev(g->nodes, g->sels->buffer[i]);
			node[i * 5 + 0] = u->base; node[i * 5 + 1] = u->bonus; node[i * 5 + 2] = u->rpos; node[i * 5 + 3] = u->nct; node[i * 5 + 4] = u->mmidx;
			eoff[i] = nedge;
			for(eidx=u->edge;eidx;eidx=ref_bspoaedgev(g->edges, eidx)->next) if(get_bitvec(g->states, ref_bspoaedgev(g->edges, eidx)->node)) nedge ++;
		}
		eoff[nloc] = nedge;
		edst = malloc(sizeof(int32_t) * (nedge + 1)); nedge = 0;
		for(i=0;i<nloc;i++){
			bspoanode_t *u = ref_bspoanodev(g->nodes, g->sels->buffer[i]);
			for(eidx=u->edge;eidx;eidx=ref_bspoaedgev(g->edges, eidx)->next){
				bspoaedge_t *e = ref_bspoaedgev(g->edges, eidx);
				if(get_bitvec(g->states, e->node)) edst[nedge ++] = loc[e->node];
			}
		}
		memset(hdr, 0, sizeof(hdr));
		hdr[0] = 0x504F4131; hdr[1] = bw; hdr[2] = g->piecewise; hdr[3] = g->slen; hdr[4] = par->alnmode; hdr[5] = par->M; hdr[6] = par->X;
		hdr[7] = par->O; hdr[8] = par->E; hdr[9] = par->Q; hdr[10] = par->P; hdr[11] = par->T; hdr[12] = par->refbonus; hdr[13] = nloc;
		hdr[14] = loc[nhead]; hdr[15] = loc[ntail]; hdr[16] = nedge; hdr[20] = (int32_t)g->mmblk; hdr[23] = g->qb;
		*hdr_at = blob->n;
		blob_put(blob, hdr, sizeof(hdr));
		blob_put(blob, g->qseq->buffer + g->qb, g->slen);
		blob_put(blob, node, sizeof(int32_t) * 5 * nloc);
		blob_put(blob, eoff, sizeof(int32_t) * (nloc + 1));
		blob_put(blob, edst, sizeof(int32_t) * nedge);
		free(node); free(eoff); free(edst);
		{	/* reverse edges (u->erev lists, bspoa.h:2319, 2429) restricted to selected nodes, with their coverage: what alignment2graph_bspoa walks */
			u4i nre = 0;
			int32_t *reoff = malloc(sizeof(int32_t) * (nloc + 1)), *resrc, *recov;
			for(i=0;i<nloc;i++){
				bspoanode_t *u = ref_bspoanodev(g->nodes, g->sels->buffer[i]);
				reoff[i] = nre;
				for(eidx=u->erev;eidx;eidx=ref_bspoaedgev(g->edges, eidx)->next) if(get_bitvec(g->states, ref_bspoaedgev(g->edges, eidx)->node)) nre ++;
			}
			reoff[nloc] = nre;
			resrc = malloc(sizeof(int32_t) * (nre + 1)); recov = malloc(sizeof(int32_t) * (nre + 1)); nre = 0;
			for(i=0;i<nloc;i++){
				bspoanode_t *u = ref_bspoanodev(g->nodes, g->sels->buffer[i]);
				for(eidx=u->erev;eidx;eidx=ref_bspoaedgev(g->edges, eidx)->next){
					bspoaedge_t *e = ref_bspoaedgev(g->edges, eidx);
					if(get_bitvec(g->states, e->node)){ resrc[nre] = loc[e->node]; recov[nre] = e->cov; nre ++; }
				}
			}
			((int32_t*)(blob->buf + *hdr_at))[21] = nre;
			blob_put(blob, reoff, sizeof(int32_t) * (nloc + 1));
			blob_put(blob, resrc, sizeof(int32_t) * nre);
			blob_put(blob, recov, sizeof(int32_t) * nre);
			free(reoff); free(resrc); free(recov);
		}
	} else {
		int32_t *hdr = (int32_t*)(blob->buf + *hdr_at);
		int8_t *row = malloc(3 * (size_t)bw); int32_t ub[17]; uint8_t *done = malloc(nloc + 4);
		hdr[17] = g->maxscr; hdr[18] = g->maxidx >= 0 ? (int32_t)loc[g->maxidx] : -1; hdr[19] = g->maxoff;
		for(i=0;i<nloc;i++){
			bspoanode_t *u = ref_bspoanodev(g->nodes, g->sels->buffer[i]);
			b1i *us, *es, *qs; int *ubegs;
			dpalign_row_prepare_data(g, u->mmidx, &us, &es, &qs, &ubegs);
			memset(row, 0, 3 * (size_t)bw);
			for(p=0;p<bw;p++){
				u4i idx = (p % W) * WORDSIZE + (p / W);
				row[p] = us[idx];
				if(es) row[bw + p] = es[idx];
				if(qs) row[2 * bw + p] = qs[idx];
			}
			memcpy(ub, ubegs, sizeof(ub));
			blob_put(blob, row, 3 * (size_t)bw);
			blob_put(blob, ub, sizeof(ub));
			done[i] = (offset_bspoanodev(g->nodes, u) == nhead) ? 1 : (u->vst ? 1 : 0);
		}
		blob_put(blob, done, nloc);
		free(row); free(done);
	}
}

/* reads: one base per byte (0..3).  par_override: NULL, or 10 ints {bandwidth, M, X, O, E, Q, P, T, refbonus, alnmode}.
 * Returns the number of sweep jobs dumped; *out receives a malloc'ed blob (free with bsref_free). */
int64_t bsref_poa_dump(uint32_t nreads, const uint8_t *seqs, const uint64_t *off, const uint32_t *len, const int32_t *par_override,
		uint8_t **out, uint64_t *out_len){
	BSPOAPar par = DEFAULT_BSPOA_PAR;
	BSPOA *g;
	bsref_blob_t blob = {NULL, 0, 0};
	int64_t njobs = 0;
	u4i i; u2i rid;
	if(par_override){
		par.bandwidth = par_override[0]; par.M = par_override[1]; par.X = par_override[2]; par.O = par_override[3]; par.E = par_override[4];
		par.Q = par_override[5]; par.P = par_override[6]; par.T = par_override[7]; par.refbonus = par_override[8]; par.alnmode = par_override[9];
	}
	par.realn = 0;
	g = init_bspoa(par);
	beg_bspoa(g);
	for(i=0;i<nreads;i++) fwdbitseqpush_bspoa(g, (u1i*)seqs + off[i], len[i]);
	/* end_bspoa, bspoa.h:4737-4760, with the dump hooks */
	if(g->seqs->nseq > 1){
		if(g->par->shuffle) shuffle_reads_by_kmers_bspoa(g);
		g->nmsa = g->par->seqcore ? num_min(g->seqs->nseq, g->par->seqcore) : g->seqs->nseq;
		for(rid=0;rid<g->seqs->nseq;rid++) _add_read_bspoa_core(g, rid);
		g->nrds = 1;
		for(rid=1;rid<g->nmsa;rid++){
			u4i nhead, ntail, rlen = g->seqs->rdlens->buffer[rid];
			u2i ridxbeg, ridxend;
			size_t hdr_at = 0;
			int score; seqalign_result_t rs;
			if(!g->par->refmode && g->par->bwtrigger){ msa_bspoa(g); simple_cns_bspoa(g); }
			/* align_rd_bspoa(g, par, 0, rid, 0, rlen), bspoa.h:2620-2650 */
			clear_u8v(g->todels);
			if(rlen){
				nhead = get_rdnode_bspoa(g, rid, -1)->header;
				ntail = get_rdnode_bspoa(g, rid, rlen)->header;
				if(g->par->nrec){ ridxbeg = num_max(0, Int(rid) - g->par->nrec - 1); ridxend = rid; } else { ridxbeg = 0; ridxend = MAX_U2; }
				sel_nodes_bspoa(g, nhead, ntail, ridxbeg, ridxend);
				prepare_rd_align_bspoa(g, g->par, nhead, ntail, rid, 0, rlen);
				bsref_poa_dump_job(g, g->par, nhead, ntail, &blob, 0, &hdr_at);
				score = align_rd_bspoacore(g, g->par, rid, nhead, ntail);
				bsref_poa_dump_job(g, g->par, nhead, ntail, &blob, 1, &hdr_at);
				njobs ++;
				rs = alignment2graph_bspoa(g, g->par, rid, 0, nhead, ntail, g->maxidx, g->maxoff, NULL);
				UNUSED(score);
				{	/* what the traceback decided: the result counts and, per read position, the graph node it was merged into (-1: none) */
					int32_t r10[10], *mh = malloc(sizeof(int32_t) * (g->slen + 1));
					u4i x, k, nloc2 = g->sels->size;
					u4i *loc2 = malloc(sizeof(u4i) * (g->nodes->size + 1));
					for(k=0;k<nloc2;k++) loc2[g->sels->buffer[k]] = k;
					bsref_store_result(r10, &rs);
					for(x=0;x<g->slen;x++){
						bspoanode_t *rn = get_rdnode_bspoa(g, rid, g->qb + x);
						mh[x] = (rn->header < g->nodes->size && get_bitvec(g->states, rn->header)) ? (int32_t)loc2[rn->header] : -1;
					}
					blob_put(&blob, r10, sizeof(r10));
					blob_put(&blob, mh, sizeof(int32_t) * g->slen);
					free(mh); free(loc2);
				}
				for(i=0;i<g->todels->size;i++){
					chg_edge_bspoa(g, ref_bspoanodev(g->nodes, g->todels->buffer[i] >> 32), ref_bspoanodev(g->nodes, g->todels->buffer[i] & MAX_U4), -1, NULL);
				}
				clear_u8v(g->todels);
			}
			g->nrds ++;
		}
	}
	free_bspoa(g);
	*out = blob.buf; *out_len = blob.n;
	return njobs;
}

void bsref_free(void *p){ free(p); }

/* a whole BSPOA job (DEFAULT_BSPOA_PAR) through the unmodified end_bspoa, then the reference's own dump_binary_msa_bspoa (bspoa.h:1555) into
   a malloc'ed buffer: the golden bytes of the binary MSA format */
int64_t bsref_poa_msa_bytes(uint32_t nreads, const uint8_t *seqs, const uint64_t *off, const uint32_t *len, const char *meta, uint32_t metalen, uint8_t **out){
	BSPOAPar par = DEFAULT_BSPOA_PAR;
	BSPOA *g = init_bspoa(par);
	char *buf = NULL; size_t sz = 0;
	FILE *f;
	u4i i;
	beg_bspoa(g);
	for(i=0;i<nreads;i++) fwdbitseqpush_bspoa(g, (u1i*)seqs + off[i], len[i]);
	end_bspoa(g);
	f = open_memstream(&buf, &sz);
	dump_binary_msa_bspoa(g, (char*)meta, metalen, f);
	fclose(f);
	free_bspoa(g);
	*out = (uint8_t*)buf;
	return (int64_t)sz;
}

/* Time the reference's own sweep: run a whole POA job (same call sequence as end_bspoa, bspoa.h:4737-4760, realn = 0) and
 * accumulate the CPU time spent inside align_rd_bspoacore only (thread CPU clock), plus the number of row updates (every update
 * increments v->vst once, bspoa.h:2592) and merges.  One BSPOA per call: callers may run several calls on different threads. */
#include <time.h>
static double bsref_thread_seconds(void){ struct timespec ts; clock_gettime(CLOCK_THREAD_CPUTIME_ID, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

int bsref_poa_time(uint32_t nreads, const uint8_t *seqs, const uint64_t *off, const uint32_t *len, const int32_t *par_override,
		double *dp_seconds, double *total_seconds, uint64_t *nupd_out, uint64_t *nmrg_out, uint64_t *cells_out){
	BSPOAPar par = DEFAULT_BSPOA_PAR;
	BSPOA *g;
	u4i i; u2i rid;
	double dp = 0, t0, t1, tb = bsref_thread_seconds();
	uint64_t nupd = 0, nmrg = 0, cells = 0;
	if(par_override){
		par.bandwidth = par_override[0]; par.M = par_override[1]; par.X = par_override[2]; par.O = par_override[3]; par.E = par_override[4];
		par.Q = par_override[5]; par.P = par_override[6]; par.T = par_override[7]; par.refbonus = par_override[8]; par.alnmode = par_override[9];
	}
	par.realn = 0;
	g = init_bspoa(par);
	beg_bspoa(g);
	for(i=0;i<nreads;i++) fwdbitseqpush_bspoa(g, (u1i*)seqs + off[i], len[i]);
	if(g->seqs->nseq > 1){
		if(g->par->shuffle) shuffle_reads_by_kmers_bspoa(g);
		g->nmsa = g->par->seqcore ? num_min(g->seqs->nseq, g->par->seqcore) : g->seqs->nseq;
		for(rid=0;rid<g->seqs->nseq;rid++) _add_read_bspoa_core(g, rid);
		g->nrds = 1;
		for(rid=1;rid<g->nmsa;rid++){
			u4i nhead, ntail, rlen = g->seqs->rdlens->buffer[rid], k;
			u2i ridxbeg, ridxend;
			if(!g->par->refmode && g->par->bwtrigger){ msa_bspoa(g); simple_cns_bspoa(g); }
			clear_u8v(g->todels);
			if(rlen){
				nhead = get_rdnode_bspoa(g, rid, -1)->header;
				ntail = get_rdnode_bspoa(g, rid, rlen)->header;
				if(g->par->nrec){ ridxbeg = num_max(0, Int(rid) - g->par->nrec - 1); ridxend = rid; } else { ridxbeg = 0; ridxend = MAX_U2; }
				sel_nodes_bspoa(g, nhead, ntail, ridxbeg, ridxend);
				prepare_rd_align_bspoa(g, g->par, nhead, ntail, rid, 0, rlen);
				t0 = bsref_thread_seconds();
				align_rd_bspoacore(g, g->par, rid, nhead, ntail);
				t1 = bsref_thread_seconds();
				dp += t1 - t0;
				for(k=0;k<g->sels->size;k++){
					bspoanode_t *u = ref_bspoanodev(g->nodes, g->sels->buffer[k]);
					if(g->sels->buffer[k] == ntail || g->sels->buffer[k] == nhead) continue;
					nupd += u->vst; if(u->vst > 1) nmrg += u->vst - 1;
					cells += (uint64_t)u->vst * g->bandwidth;
				}
				alignment2graph_bspoa(g, g->par, rid, 0, nhead, ntail, g->maxidx, g->maxoff, NULL);
				for(k=0;k<g->todels->size;k++){
					chg_edge_bspoa(g, ref_bspoanodev(g->nodes, g->todels->buffer[k] >> 32), ref_bspoanodev(g->nodes, g->todels->buffer[k] & MAX_U4), -1, NULL);
				}
				clear_u8v(g->todels);
			}
			g->nrds ++;
		}
	}
	free_bspoa(g);
	*dp_seconds = dp; *total_seconds = bsref_thread_seconds() - tb;
	*nupd_out = nupd; *nmrg_out = nmrg; *cells_out = cells;
	return 0;
}
