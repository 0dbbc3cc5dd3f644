/*
 * bsalign_b200_poa_compat.h -- GNU-C drop-in for the read-vs-graph DP sweep of ruanjue/bsalign's BSPOA, re-bodied on top of
 * libbsalign_b200.so (C ABI: bsalign_b200.h, bsb200_poa_*).  Include it AFTER the reference's own bspoa.h.
 *
 *   b200_align_rd_bspoacore(ctx, g, par, rid, nhead, ntail)
 *       same contract as align_rd_bspoacore (bspoa.h:2515-2618) for ONE BSPOA: after it returns, every selected node's row
 *       block sits in g->memp (dpalign_row_prepare_data, bspoa.h:1787-1793) and g->maxscr / g->maxidx / g->maxoff are set,
 *       so alignment2graph_bspoa (bspoa.h:2274) runs unchanged.  One job per call pays a PCIe round trip.
 *
 *   b200_end_bspoa_batch(ctx, gs, n)
 *       end_bspoa (bspoa.h:4722-4778) for n BSPOA objects stepped in LOCK-STEP: in round r every object prepares the alignment of
 *       its read r on the host exactly as the reference does (msa_bspoa, simple_cns_bspoa, sel_nodes_bspoa,
 *       prepare_rd_align_bspoa), the n sweeps AND the walk of alignment2graph_bspoa (bspoa.h:2274-2497) run as ONE GPU batch, and
 *       every object then replays the graph surgery of alignment2graph_bspoa (merge_nodes_bspoa, cpos, connect_rdnode_bspoa) from
 *       the walk's decisions, in the reference's order.  Only 4 bytes per read position come back from the GPU.
 *       With -DBSALIGN_B200_POA_HOST_TRACEBACK the row blocks are copied into g->memp instead and the reference's own
 *       alignment2graph_bspoa runs on them (15 MB per sweep of a 15 kb read).
 *
 * -DBSALIGN_B200_OVERRIDE routes the reference name end_bspoa to the batch driver with n = 1.
 * Programmer errors keep the reference's behaviour: message on stderr + abort().  There is no CPU fallback.
 */
#ifndef BSALIGN_B200_POA_COMPAT_H
#define BSALIGN_B200_POA_COMPAT_H

#include "bsalign_b200.h"

#ifndef BANDED_STRIPED_SIMD_RECURRENT_ALIGNMENT_GRAPH_MSA_CNS_RJ_H
#error "include the reference's bspoa.h before bsalign_b200_poa_compat.h"
#endif

/* packed arenas of a batch of sweep jobs: the argument layout of bsb200_poa_rows_batch */
typedef struct {
	uint32_t njobs, cap_jobs;
	int32_t *par; uint64_t *qoff; uint32_t *slen; uint64_t *node_off, *edge_off; uint32_t *head, *tail;
	uint8_t *queries; uint64_t nq, cap_q;
	uint8_t *base, *bonus; int32_t *rpos, *nct; uint64_t nnode, cap_node;
	int32_t *eoff; uint64_t neoff, cap_eoff;
	int32_t *edst; uint64_t nedge, cap_edge;
	uint32_t *loc; uint64_t cap_loc;          /* global node id -> local id scratch */
	uint8_t *rows; uint64_t cap_rows; uint64_t *row_off;
	int32_t *best, *status;
	/* reverse edges + what the device-side walk returns */
	int32_t *reoff; uint64_t nreoff, cap_reoff;
	int32_t *resrc, *recov; uint64_t nredge, cap_redge; uint64_t *redge_off;
	int32_t *match; uint64_t cap_match; int32_t *trace;
} b200_poa_pack_t;

#define B200_GROW(ptr, cap, need, type) do { if((uint64_t)(need) > (cap)){ (cap) = (uint64_t)(need) * 3 / 2 + 64; (ptr) = (type*)realloc((ptr), (cap) * sizeof(type)); \
	if((ptr) == NULL){ fflush(stdout); fprintf(stderr, " -- Out of memory in %s -- %s:%d --\n", __FUNCTION__, __FILE__, __LINE__); fflush(stderr); abort(); } } } while(0)

static inline b200_poa_pack_t* b200_poa_pack_init(void){ return (b200_poa_pack_t*)calloc(1, sizeof(b200_poa_pack_t)); }
static inline void b200_poa_pack_clear(b200_poa_pack_t *p){ p->njobs = 0; p->nq = 0; p->nnode = 0; p->neoff = 0; p->nedge = 0; p->nreoff = 0; p->nredge = 0; }
static inline void b200_poa_pack_free(b200_poa_pack_t *p){
	free(p->par); free(p->qoff); free(p->slen); free(p->node_off); free(p->edge_off); free(p->head); free(p->tail); free(p->queries);
	free(p->base); free(p->bonus); free(p->rpos); free(p->nct); free(p->eoff); free(p->edst); free(p->loc); free(p->rows); free(p->row_off);
	free(p->best); free(p->status); free(p->reoff); free(p->resrc); free(p->recov); free(p->redge_off); free(p->match); free(p->trace); free(p);
}

/* append the sweep job of g (after prepare_rd_align_bspoa) to the pack: what align_rd_bspoacore would consume */
static inline void b200_poa_pack_job(b200_poa_pack_t *p, BSPOA *g, BSPOAPar *par, u4i nhead, u4i ntail){
	u4i i, nloc = g->sels->size, eidx, j = p->njobs;
	uint64_t dummy_cap;
	if(j + 2 > p->cap_jobs){
		uint32_t nc = (j + 2) * 2 + 14;
		p->par = (int32_t*)realloc(p->par, sizeof(int32_t) * 10 * nc); p->qoff = (uint64_t*)realloc(p->qoff, 8 * (size_t)nc); p->slen = (uint32_t*)realloc(p->slen, 4 * (size_t)nc);
		p->node_off = (uint64_t*)realloc(p->node_off, 8 * (size_t)nc); p->edge_off = (uint64_t*)realloc(p->edge_off, 8 * (size_t)nc);
		p->head = (uint32_t*)realloc(p->head, 4 * (size_t)nc); p->tail = (uint32_t*)realloc(p->tail, 4 * (size_t)nc);
		p->row_off = (uint64_t*)realloc(p->row_off, 8 * (size_t)nc); p->best = (int32_t*)realloc(p->best, 12 * (size_t)nc); p->status = (int32_t*)realloc(p->status, 4 * (size_t)nc);
		p->redge_off = (uint64_t*)realloc(p->redge_off, 8 * (size_t)nc); p->trace = (int32_t*)realloc(p->trace, 32 * (size_t)nc);
		p->cap_jobs = nc;
	}
	B200_GROW(p->loc, p->cap_loc, g->nodes->size + 1, uint32_t);
	for(i=0;i<nloc;i++) p->loc[g->sels->buffer[i]] = i;
	{
		int32_t *q = p->par + 10 * (size_t)j;
		q[0] = g->bandwidth; q[1] = par->alnmode; q[2] = par->M; q[3] = par->X; q[4] = par->O; q[5] = par->E; q[6] = par->Q; q[7] = par->P; q[8] = par->T; q[9] = par->refbonus;
	}
	p->qoff[j] = p->nq; p->slen[j] = g->slen;
	B200_GROW(p->queries, p->cap_q, p->nq + g->slen + 16, uint8_t);
	memcpy(p->queries + p->nq, g->qseq->buffer + g->qb, g->slen); p->nq += g->slen;
	p->node_off[j] = p->nnode; p->edge_off[j] = p->nedge;
	dummy_cap = p->cap_node; B200_GROW(p->base, dummy_cap, p->nnode + nloc, uint8_t);
	dummy_cap = p->cap_node; B200_GROW(p->bonus, dummy_cap, p->nnode + nloc, uint8_t);
	dummy_cap = p->cap_node; B200_GROW(p->rpos, dummy_cap, p->nnode + nloc, int32_t);
	B200_GROW(p->nct, p->cap_node, p->nnode + nloc, int32_t);
	B200_GROW(p->eoff, p->cap_eoff, p->neoff + nloc + 1, int32_t);
	{
		int32_t *eo = p->eoff + p->neoff;
		uint64_t ne = 0;
		for(i=0;i<nloc;i++){
			bspoanode_t *u = ref_bspoanodev(g->nodes, g->sels->buffer[i]);
			uint64_t k = p->nnode + i;
			p->base[k] = u->base; p->bonus[k] = u->bonus; p->rpos[k] = u->rpos; p->nct[k] = u->nct;
			eo[i] = (int32_t)ne;
			for(eidx=u->edge;eidx;eidx=ref_bspoaedgev(g->edges, eidx)->next){   /* the edge-list order of bspoa.h:2533-2538 */
				bspoaedge_t *e = ref_bspoaedgev(g->edges, eidx);
				if(get_bitvec(g->states, e->node) == 0) continue;
				B200_GROW(p->edst, p->cap_edge, p->nedge + ne + 1, int32_t);
				p->edst[p->nedge + ne] = (int32_t)p->loc[e->node];
				ne ++;
			}
		}
		eo[nloc] = (int32_t)ne;
		p->nedge += ne;
	}
	/* reverse edges: the erev lists alignment2graph_bspoa walks (bspoa.h:2319, 2429), with their coverage (bspoa.h:2457-2464) */
	p->redge_off[j] = p->nredge;
	B200_GROW(p->reoff, p->cap_reoff, p->nreoff + nloc + 1, int32_t);
	{
		int32_t *ro = p->reoff + p->nreoff;
		uint64_t ne = 0;
		for(i=0;i<nloc;i++){
			bspoanode_t *u = ref_bspoanodev(g->nodes, g->sels->buffer[i]);
			ro[i] = (int32_t)ne;
			for(eidx=u->erev;eidx;eidx=ref_bspoaedgev(g->edges, eidx)->next){
				bspoaedge_t *e = ref_bspoaedgev(g->edges, eidx);
				if(get_bitvec(g->states, e->node) == 0) continue;
				dummy_cap = p->cap_redge; B200_GROW(p->resrc, dummy_cap, p->nredge + ne + 1, int32_t);
				B200_GROW(p->recov, p->cap_redge, p->nredge + ne + 1, int32_t);
				p->resrc[p->nredge + ne] = (int32_t)p->loc[e->node];
				p->recov[p->nredge + ne] = (int32_t)e->cov;
				ne ++;
			}
		}
		ro[nloc] = (int32_t)ne;
		p->nredge += ne;
	}
	p->nreoff += nloc + 1;
	p->redge_off[j + 1] = p->nredge;
	p->neoff += nloc + 1; p->nnode += nloc;
	p->head[j] = p->loc[nhead]; p->tail[j] = p->loc[ntail];
	p->njobs = j + 1;
	p->node_off[j + 1] = p->nnode; p->edge_off[j + 1] = p->nedge;
}

/* append the ONE job of q (packed by its object's host thread) to the batch p: plain copies, every index inside a job is local to it */
static inline void b200_poa_pack_append(b200_poa_pack_t *p, const b200_poa_pack_t *q){
	u4i j = p->njobs;
	uint64_t nloc = q->nnode, ne = q->nedge, nre = q->nredge, dummy_cap;
	if(j + 2 > p->cap_jobs){
		uint32_t nc = (j + 2) * 2 + 14;
		p->par = (int32_t*)realloc(p->par, sizeof(int32_t) * 10 * nc); p->qoff = (uint64_t*)realloc(p->qoff, 8 * (size_t)nc); p->slen = (uint32_t*)realloc(p->slen, 4 * (size_t)nc);
		p->node_off = (uint64_t*)realloc(p->node_off, 8 * (size_t)nc); p->edge_off = (uint64_t*)realloc(p->edge_off, 8 * (size_t)nc);
		p->head = (uint32_t*)realloc(p->head, 4 * (size_t)nc); p->tail = (uint32_t*)realloc(p->tail, 4 * (size_t)nc);
		p->row_off = (uint64_t*)realloc(p->row_off, 8 * (size_t)nc); p->best = (int32_t*)realloc(p->best, 12 * (size_t)nc); p->status = (int32_t*)realloc(p->status, 4 * (size_t)nc);
		p->redge_off = (uint64_t*)realloc(p->redge_off, 8 * (size_t)nc); p->trace = (int32_t*)realloc(p->trace, 32 * (size_t)nc);
		p->cap_jobs = nc;
	}
	memcpy(p->par + 10 * (size_t)j, q->par, sizeof(int32_t) * 10);
	p->qoff[j] = p->nq; p->slen[j] = q->slen[0]; p->head[j] = q->head[0]; p->tail[j] = q->tail[0];
	B200_GROW(p->queries, p->cap_q, p->nq + q->nq + 16, uint8_t);
	memcpy(p->queries + p->nq, q->queries, q->nq); p->nq += q->nq;
	p->node_off[j] = p->nnode; p->edge_off[j] = p->nedge; p->redge_off[j] = p->nredge;
	dummy_cap = p->cap_node; B200_GROW(p->base, dummy_cap, p->nnode + nloc, uint8_t);
	dummy_cap = p->cap_node; B200_GROW(p->bonus, dummy_cap, p->nnode + nloc, uint8_t);
	dummy_cap = p->cap_node; B200_GROW(p->rpos, dummy_cap, p->nnode + nloc, int32_t);
	B200_GROW(p->nct, p->cap_node, p->nnode + nloc, int32_t);
	memcpy(p->base + p->nnode, q->base, nloc); memcpy(p->bonus + p->nnode, q->bonus, nloc);
	memcpy(p->rpos + p->nnode, q->rpos, sizeof(int32_t) * nloc); memcpy(p->nct + p->nnode, q->nct, sizeof(int32_t) * nloc);
	B200_GROW(p->eoff, p->cap_eoff, p->neoff + nloc + 1, int32_t);
	memcpy(p->eoff + p->neoff, q->eoff, sizeof(int32_t) * (nloc + 1));
	B200_GROW(p->edst, p->cap_edge, p->nedge + ne + 1, int32_t);
	memcpy(p->edst + p->nedge, q->edst, sizeof(int32_t) * ne);
	B200_GROW(p->reoff, p->cap_reoff, p->nreoff + nloc + 1, int32_t);
	memcpy(p->reoff + p->nreoff, q->reoff, sizeof(int32_t) * (nloc + 1));
	dummy_cap = p->cap_redge; B200_GROW(p->resrc, dummy_cap, p->nredge + nre + 1, int32_t);
	B200_GROW(p->recov, p->cap_redge, p->nredge + nre + 1, int32_t);
	memcpy(p->resrc + p->nredge, q->resrc, sizeof(int32_t) * nre); memcpy(p->recov + p->nredge, q->recov, sizeof(int32_t) * nre);
	p->nnode += nloc; p->neoff += nloc + 1; p->nreoff += nloc + 1; p->nedge += ne; p->nredge += nre;
	p->njobs = j + 1;
	p->node_off[j + 1] = p->nnode; p->edge_off[j + 1] = p->nedge; p->redge_off[j + 1] = p->nredge;
}

/* run every packed sweep on the GPU (one batch) */
static inline void b200_poa_pack_run(bsb200_ctx *ctx, b200_poa_pack_t *p){
	uint32_t j;
	uint64_t tot = 0;
	if(p->njobs == 0) return;
	for(j=0;j<p->njobs;j++){ p->row_off[j] = tot; tot += (p->node_off[j + 1] - p->node_off[j]) * (uint64_t)bsb200_poa_block_bytes(p->par + 10 * (size_t)j); }
	p->row_off[p->njobs] = tot;
	B200_GROW(p->rows, p->cap_rows, tot + 16, uint8_t);
	if(bsb200_poa_rows_batch(ctx, p->njobs, p->par, p->queries, p->qoff, p->slen, p->node_off, p->base, p->bonus, p->rpos, p->nct, p->eoff, p->edge_off, p->edst,
			p->head, p->tail, p->rows, p->best, p->status, NULL)){
		fflush(stdout); fprintf(stderr, " -- bsalign_b200: %s in %s -- %s:%d --\n", bsb200_last_error(ctx), __FUNCTION__, __FILE__, __LINE__); fflush(stderr);
		abort();
	}
}

/* run every packed sweep AND the walk of alignment2graph_bspoa on the GPU; the row blocks stay in HBM */
static inline void b200_poa_pack_run_walk(bsb200_ctx *ctx, b200_poa_pack_t *p){
	if(p->njobs == 0) return;
	B200_GROW(p->match, p->cap_match, p->nq + 16, int32_t);
	if(bsb200_poa_align_batch(ctx, p->njobs, p->par, p->queries, p->qoff, p->slen, p->node_off, p->base, p->bonus, p->rpos, p->nct, p->eoff, p->edge_off, p->edst,
			p->head, p->tail, p->reoff, p->redge_off, p->resrc, p->recov, NULL, p->best, p->status, NULL, p->match, p->trace)){
		fflush(stdout); fprintf(stderr, " -- bsalign_b200: %s in %s -- %s:%d --\n", bsb200_last_error(ctx), __FUNCTION__, __FILE__, __LINE__); fflush(stderr);
		abort();
	}
}

/* The side effects of alignment2graph_bspoa (bspoa.h:2274-2512) replayed from the device walk of job j: per read position the node it
 * was aligned to.  merge_nodes_bspoa / cpos in decreasing read position like the reference's walk, then its closing loop. */
static inline void b200_poa_pack_replay(b200_poa_pack_t *p, uint32_t j, BSPOA *g, u4i rid, u4i rbeg, u4i nhead, u4i ntail){
	const int32_t *match = p->match + p->qoff[j], *tr = p->trace + 8 * (size_t)j;
	bspoanode_t *u, *v, *n;
	int x, cpos;
	g->maxscr = p->best[3 * (size_t)j];
	g->maxidx = p->best[3 * (size_t)j + 1] >= 0 ? (int)g->sels->buffer[p->best[3 * (size_t)j + 1]] : -1;
	g->maxoff = p->best[3 * (size_t)j + 2];
	if(tr[7] || g->maxidx < 0){   /* the reference itself reads outside a row or never leaves its walk on this input */
		fflush(stdout); fprintf(stderr, " -- bsalign_b200: inconsistent traceback (flags %d) in %s -- %s:%d --\n", tr[7], __FUNCTION__, __FILE__, __LINE__); fflush(stderr);
		abort();
	}
	nhead = ref_bspoanodev(g->nodes, nhead)->header;
	ntail = ref_bspoanodev(g->nodes, ntail)->header;
	u = get_rdnode_bspoa(g, rid, 0);
	v = get_rdnode_bspoa(g, rid, g->qlen);
	while(u < v){ u->cpos = 0; u ++; }
	cpos = ref_bspoanodev(g->nodes, g->maxidx)->cpos;
	for(x=g->maxoff;x>tr[0]&&x>=0;x--){
		if(match[x] < 0) continue;                                   /* insertion */
		n = ref_bspoanodev(g->nodes, g->sels->buffer[match[x]]);
		u = get_rdnode_bspoa(g, rid, rbeg + g->qb + x);
		u->cpos = n->cpos;                                             /* bspoa.h:2394-2395 */
		if(offset_bspoanodev(g->nodes, n) != nhead && offset_bspoanodev(g->nodes, n) != ntail && u->base == n->base){
			merge_nodes_bspoa(g, n, u);                                /* bspoa.h:2402-2404 */
		}
	}
	v = get_rdnode_bspoa(g, rid, rbeg + g->qlen);                      /* bspoa.h:2500-2511 */
	connect_rdnode_bspoa(g, rid, rbeg + g->qlen);
	for(x=g->qlen-1;x>=0;x--){
		connect_rdnode_bspoa(g, rid, rbeg + x);
		v --;
		if(v->cpos){
			cpos = v->cpos;
		} else if(x < Int(g->qlen)){
			v->cpos = cpos;
		}
	}
}

/* hand job j's results to its BSPOA: row blocks into g->memp from block 2 on (bspoa.h:2218-2223), best end into g->max* */
static inline int b200_poa_pack_take(b200_poa_pack_t *p, uint32_t j, BSPOA *g){
	uint64_t bytes = p->row_off[j + 1] - p->row_off[j];
	memcpy(g->memp->buffer + 2 * g->mmblk, p->rows + p->row_off[j], bytes);
	g->maxscr = p->best[3 * (size_t)j];
	g->maxidx = p->best[3 * (size_t)j + 1] >= 0 ? (int)g->sels->buffer[p->best[3 * (size_t)j + 1]] : -1;
	g->maxoff = p->best[3 * (size_t)j + 2];
	return g->maxscr;
}

static inline int b200_align_rd_bspoacore(bsb200_ctx *ctx, BSPOA *g, BSPOAPar *par, u2i rid, u4i nhead, u4i ntail){
	b200_poa_pack_t *p;
	int scr;
	UNUSED(rid);
	if(g->sels->size == 0) return align_rd_bspoacore(g, par, rid, nhead, ntail);   /* nothing selected: no DP rows at all */
	p = b200_poa_pack_init();
	b200_poa_pack_job(p, g, par, nhead, ntail);
	b200_poa_pack_run(ctx, p);
	scr = b200_poa_pack_take(p, 0, g);
	b200_poa_pack_free(p);
	return scr;
}

/* end_bspoa (bspoa.h:4722-4778) for n objects in lock-step; the read-vs-graph sweeps of one round are one GPU batch.  The host work
 * around the sweeps (graph surgery, msa, consensus, re-alignment: the reference's own code) is per object and independent: compiled
 * b200_poa_host_threads sets how many host threads share it (one per in-flight object). */
#include <pthread.h>
#include <unistd.h>
#include <time.h>
typedef struct {
	bsb200_ctx *ctx; BSPOA **gs; u4i n, *nheads, *ntails, *slot; u2i rid; b200_poa_pack_t *p;
	int phase; volatile u4i next;
	b200_poa_pack_t **mine;   /* per object: its sweep job of this round, packed by the object's own host thread */
#ifdef BSALIGN_B200_POA_KMER_H
	b200_poa_kmer_slot_t *kslot;   /* per object: the band-placement alignment of this round, computed as one GPU batch */
#endif
} b200_poa_round_t;

static inline void b200_poa_round_object(b200_poa_round_t *r, u4i k){
	BSPOA *g = r->gs[k];
	u2i rid = r->rid;
	if(r->phase == 4){   /* bspoa.h:4726-4751: the reads into the graph */
		clear_u1v(g->cns); clear_u1v(g->qlt); clear_u1v(g->alt);
		if(g->par->refmode){
			resize_u1v(g->cns, g->seqs->rdlens->buffer[0]); resize_u1v(g->qlt, g->seqs->rdlens->buffer[0]); resize_u1v(g->alt, g->seqs->rdlens->buffer[0]);
			bitseq_basebank(g->seqs->rdseqs, g->seqs->rdoffs->buffer[0], g->seqs->rdlens->buffer[0], g->cns->buffer);
			memset(g->qlt->buffer, 0, g->seqs->rdlens->buffer[0]); memset(g->alt->buffer, 0, g->seqs->rdlens->buffer[0]);
		}
		if(g->seqs->nseq <= 1){ g->nmsa = 0; return; }
		if(g->par->shuffle) shuffle_reads_by_kmers_bspoa(g);
		g->nmsa = g->par->seqcore? num_min(g->seqs->nseq, g->par->seqcore) : g->seqs->nseq;
		for(rid=0;rid<g->seqs->nseq;rid++) _add_read_bspoa_core(g, rid);
		g->nrds = 1;
		return;
	}
	if(r->phase == 0){   /* before the sweep: bspoa.h:4753-4755 and the head of align_rd_bspoa (bspoa.h:2620-2642) */
		u4i rlen; u2i ridxbeg, ridxend;
		r->slot[k] = MAX_U4;
		if(g->seqs->nseq <= 1 || rid >= g->nmsa) return;
		if(!g->par->refmode && g->par->bwtrigger){ msa_bspoa(g); simple_cns_bspoa(g); }
		rlen = g->seqs->rdlens->buffer[rid];
		clear_u8v(g->todels);
		if(rlen == 0){ g->nrds ++; return; }
		r->nheads[k] = get_rdnode_bspoa(g, rid, -1)->header;
		r->ntails[k] = get_rdnode_bspoa(g, rid, rlen)->header;
		if(g->par->nrec){ ridxbeg = num_max(0, Int(rid) - g->par->nrec - 1); ridxend = rid; } else { ridxbeg = 0; ridxend = MAX_U2; }
		sel_nodes_bspoa(g, r->nheads[k], r->ntails[k], ridxbeg, ridxend);
#ifdef BSALIGN_B200_POA_KMER_H
		/* with bsalign_b200_poa_kmer.h the round stops here: does prepare_rd_align_bspoa align the read against the consensus with the
		 * k-mer guided edit (the conditions of bspoa.h:2052-2088)?  Then the read is unpacked now (prepare does the same again) and the
		 * alignment of all such objects is one GPU batch between phase 0 and phase 3 */
		r->slot[k] = MAX_U4 - 3;
		r->kslot[k].armed = 0;
		if(g->par->bwtrigger && g->par->ksz && ref_bspoanodev(g->nodes, r->nheads[k])->header == g->HEAD && ref_bspoanodev(g->nodes, r->ntails[k])->header == g->TAIL
				&& !(g->par->refmode && g->cges->buffer[rid] > g->cgbs->buffer[rid]) && g->cns->size && Int(roundup_times(rlen, WORDSIZE)) > g->par->bandwidth){
			clear_and_encap_u1v(g->qseq, rlen);
			bitseq_basebank(g->seqs->rdseqs, g->seqs->rdoffs->buffer[rid], rlen, g->qseq->buffer);
			g->qseq->size = rlen;
			r->kslot[k].armed = -1;   /* wanted */
		}
	} else if(r->phase == 3){   /* the rest of phase 0 once the batch of k-mer guided alignments is back */
		u4i rlen;
		if(r->slot[k] != MAX_U4 - 3) return;
		rlen = g->seqs->rdlens->buffer[rid];
		b200_poa_kmer_slot = r->kslot[k].armed == 1 ? &r->kslot[k] : NULL;
		prepare_rd_align_bspoa(g, g->par, r->nheads[k], r->ntails[k], rid, 0, rlen);
		b200_poa_kmer_slot = NULL;
#else
		prepare_rd_align_bspoa(g, g->par, r->nheads[k], r->ntails[k], rid, 0, rlen);
#endif
		if(g->sels->size == 0){ align_rd_bspoacore(g, g->par, rid, r->nheads[k], r->ntails[k]); r->slot[k] = MAX_U4 - 1; return; }
		r->slot[k] = MAX_U4 - 2;   /* takes part in this round's batch */
		if(r->mine[k] == NULL) r->mine[k] = b200_poa_pack_init();
		b200_poa_pack_clear(r->mine[k]);
		b200_poa_pack_job(r->mine[k], g, g->par, r->nheads[k], r->ntails[k]);
	} else if(r->phase == 1){   /* behind the sweep: the tail of align_rd_bspoa (bspoa.h:2652-2666) */
		u4i t;
		if(r->slot[k] == MAX_U4) return;
#ifdef BSALIGN_B200_POA_HOST_TRACEBACK
		if(r->slot[k] != MAX_U4 - 1) b200_poa_pack_take(r->p, r->slot[k], g);
		alignment2graph_bspoa(g, g->par, rid, 0, r->nheads[k], r->ntails[k], g->maxidx, g->maxoff, NULL);
#else
		if(r->slot[k] != MAX_U4 - 1) b200_poa_pack_replay(r->p, r->slot[k], g, rid, 0, r->nheads[k], r->ntails[k]);
		else alignment2graph_bspoa(g, g->par, rid, 0, r->nheads[k], r->ntails[k], g->maxidx, g->maxoff, NULL);
#endif
		for(t=0;t<g->todels->size;t++){
			chg_edge_bspoa(g, ref_bspoanodev(g->nodes, g->todels->buffer[t] >> 32), ref_bspoanodev(g->nodes, g->todels->buffer[t] & MAX_U4), -1, NULL);
		}
		clear_u8v(g->todels);
		g->nrds ++;
	} else {   /* bspoa.h:4764-4777: the reference's own code, unchanged */
		int i;
		if(g->seqs->nseq <= 1) return;
		for(i=0;i<g->par->realn;i++){
			msa_bspoa(g);
			cns_bspoa(g);
			if(g->par->editbw < 0) remsa_edits_bspoa(g, - g->par->editbw);
			else remsa_pedits_bspoa(g, g->par->editbw / 2, 1, (i + 1 == g->par->realn));
		}
		if(g->par->shuffle) restore_rd_orders_bspoa(g);
		msa_bspoa(g);
		cns_bspoa(g);
	}
}

static void *b200_poa_round_worker(void *arg){
	b200_poa_round_t *r = (b200_poa_round_t*)arg;
	while(1){
		u4i k = __sync_fetch_and_add(&r->next, 1);
		if(k >= r->n) break;
		b200_poa_round_object(r, k);
	}
	return NULL;
}

/* objects are independent between the sweeps: one host thread per in-flight object, up to b200_poa_host_threads (0: all cores; 1: serial) */
static int b200_poa_host_threads = 1;
static inline void b200_poa_round_run(b200_poa_round_t *r, int phase){
	long nt = b200_poa_host_threads > 0 ? b200_poa_host_threads : sysconf(_SC_NPROCESSORS_ONLN);
	pthread_t th[256];
	long t;
	r->phase = phase; r->next = 0;
	if(nt > (long)r->n) nt = r->n;
	if(nt > 256) nt = 256;
	if(nt <= 1){ b200_poa_round_worker(r); return; }
	for(t=1;t<nt;t++) pthread_create(&th[t], NULL, b200_poa_round_worker, r);
	b200_poa_round_worker(r);
	for(t=1;t<nt;t++) pthread_join(th[t], NULL);
}

#ifdef BSALIGN_B200_POA_KMER_H
/* kmer_striped_seqedit_pairwise(par->ksz, g->qseq, g->cns) of every object that asked for it in phase 0 as ONE bsb200_kmer_edit_batch call;
 * the results wait in the objects' slots for the hook (bsalign_b200_poa_kmer.h).  All objects of a batch share par->ksz or are grouped by it. */
static uint32_t *b200_poa_kmer_cig = NULL; static uint64_t b200_poa_kmer_cigcap = 0;
static uint8_t *b200_poa_kmer_arena = NULL; static uint64_t b200_poa_kmer_arenacap = 0;
static unsigned long b200_poa_kmer_batches = 0, b200_poa_kmer_pairs = 0;
/* a round with fewer alignments than this keeps the reference's CPU call on the objects' host threads (a batch of 15 kb reads costs the
 * GPU ~150 ms however few pairs it holds; 64 of them cost 16 cores ~20 ms: the batch pays from a few hundred objects on).  1 = always GPU. */
static int b200_poa_kmer_min_pairs = 1;
static inline void b200_poa_kmer_round(b200_poa_round_t *r){
	u4i k, m = 0, ksz;
	uint64_t *qoff, *toff, *cgoff, bytes = 0, words = 0;
	uint32_t *qlen, *tlen, *ncg, *idx;
	bsb200_result_t *res;
	for(k=0;k<r->n;k++) if(r->slot[k] == MAX_U4 - 3 && r->kslot[k].armed == -1) m ++;
	if(m == 0) return;
	if((int)m < b200_poa_kmer_min_pairs){
		for(k=0;k<r->n;k++) if(r->slot[k] == MAX_U4 - 3 && r->kslot[k].armed == -1) r->kslot[k].armed = 0;
		return;
	}
	qoff = (uint64_t*)malloc(sizeof(uint64_t) * (3 * (size_t)m + 1)); toff = qoff + m; cgoff = toff + m;
	qlen = (uint32_t*)malloc(sizeof(uint32_t) * 4 * (size_t)m); tlen = qlen + m; ncg = tlen + m; idx = ncg + m;
	res = (bsb200_result_t*)malloc(sizeof(bsb200_result_t) * m);
	while(1){   /* one batch per distinct k-mer size (normally one) */
		u4i j = 0;
		ksz = 0; bytes = 0; words = 0;
		for(k=0;k<r->n;k++){
			BSPOA *g = r->gs[k];
			if(r->slot[k] != MAX_U4 - 3 || r->kslot[k].armed != -1) continue;
			if(ksz == 0) ksz = g->par->ksz;
			if((u4i)g->par->ksz != ksz) continue;
			idx[j] = k; qlen[j] = g->qseq->size; tlen[j] = g->cns->size;
			qoff[j] = bytes; toff[j] = bytes + qlen[j]; bytes += (uint64_t)qlen[j] + tlen[j];
			cgoff[j] = words; words += (uint64_t)qlen[j] + tlen[j] + 2;
			j ++;
		}
		if(j == 0) break;
		cgoff[j] = words;
		if(bytes > b200_poa_kmer_arenacap){ b200_poa_kmer_arenacap = bytes + bytes / 4; b200_poa_kmer_arena = (uint8_t*)realloc(b200_poa_kmer_arena, b200_poa_kmer_arenacap); }
		if(words > b200_poa_kmer_cigcap){ b200_poa_kmer_cigcap = words + words / 4; b200_poa_kmer_cig = (uint32_t*)realloc(b200_poa_kmer_cig, sizeof(uint32_t) * b200_poa_kmer_cigcap); }
		for(k=0;k<j;k++){
			BSPOA *g = r->gs[idx[k]];
			memcpy(b200_poa_kmer_arena + qoff[k], g->qseq->buffer, qlen[k]);
			memcpy(b200_poa_kmer_arena + toff[k], g->cns->buffer, tlen[k]);
		}
		if(bsb200_kmer_edit_batch(r->ctx, j, b200_poa_kmer_arena, qoff, qlen, toff, tlen, ksz, res, b200_poa_kmer_cig, cgoff, ncg, NULL)){
			fflush(stdout); fprintf(stderr, " -- bsalign_b200: %s in %s -- %s:%d --\n", bsb200_last_error(r->ctx), __FUNCTION__, __FILE__, __LINE__); fflush(stderr);
			abort();
		}
		b200_poa_kmer_batches ++; b200_poa_kmer_pairs += j;
		for(k=0;k<j;k++){
			b200_poa_kmer_slot_t *s = r->kslot + idx[k];
			s->qlen = qlen[k]; s->tlen = tlen[k];
			s->rs.score = res[k].score; s->rs.qb = res[k].qb; s->rs.qe = res[k].qe; s->rs.tb = res[k].tb; s->rs.te = res[k].te;
			s->rs.mat = res[k].mat; s->rs.mis = res[k].mis; s->rs.ins = res[k].ins; s->rs.del = res[k].del; s->rs.aln = res[k].aln;
			s->cigar = b200_poa_kmer_cig + cgoff[k]; s->ncigar = ncg[k];
			s->armed = 1;
		}
		if(j == m) break;
		/* other k-mer sizes remain: their cigars would overwrite this batch's, so those objects fall back to the reference's CPU call */
		for(k=0;k<r->n;k++) if(r->slot[k] == MAX_U4 - 3 && r->kslot[k].armed == -1) r->kslot[k].armed = 0;
		break;
	}
	free(qoff); free(qlen); free(res);
}
#endif

#ifdef BSALIGN_B200_POA_REMSA_H
/* ---- remsa_pedits_bspoa (bspoa.h:4178-4457) of all in-flight objects with its DP on the GPU -------------------------------------------
 * Phase 2 of every object (bspoa.h:4764-4777, the reference's own code) runs on its own host thread with the hook of
 * bsalign_b200_poa_remsa.h installed.  Whenever an object reaches remsa_pedit_rd_bspoacore (once per read and round) its thread puts
 * the call's inputs - the ten arrays as they lie in g->memp - into the shared arena and waits; when every object still inside phase 2
 * waits, the last one to arrive runs them as ONE bsb200_remsa_batch (the rendezvous: one job per object) and wakes the others; every
 * thread then replays the walk's merge_nodes_bspoa calls (bspoa.h:4012-4023) from the matched columns that came back, in the walk's
 * order, and returns the walk's score to remsa_pedits_bspoa, which carries on (connect_rdnodes_bspoa, next read).  Objects that finish
 * leave the rendezvous.  B200_REMSA_BATCH_FN: the batch entry point (tests substitute a CPU checker to exercise this driver without a GPU). */
#ifndef B200_REMSA_BATCH_FN
#define B200_REMSA_BATCH_FN bsb200_remsa_batch
#endif
#ifndef B200_REMSA_MAX_INFLIGHT
#define B200_REMSA_MAX_INFLIGHT 1024
#endif
typedef struct { int32_t hdr[8], out[4]; uint64_t in_off, match_off; const u1i *src[2]; uint32_t sz1; int late; } b200_remsa_job_t;
typedef struct {
	pthread_mutex_t mu; pthread_cond_t cv;
	bsb200_ctx *ctx;
	u4i active, waiting, reserved, cap_jobs; unsigned long gen;
	b200_remsa_job_t **jobs;
	uint8_t *in; uint64_t in_bytes, cap_in;
	int32_t *match; uint64_t match_ints, cap_match;
	int32_t *hdr, *out; uint64_t *in_off, *match_off;
} b200_remsa_rv_t;
static b200_remsa_rv_t *b200_remsa_rv = NULL;
static unsigned long b200_poa_remsa_batches = 0, b200_poa_remsa_jobs = 0;
/* with fewer objects than this the realn rounds keep the reference's CPU core on b200_poa_host_threads threads (a rendezvous of 20k-column
 * jobs costs ~30 ms however few objects take part; 64 such jobs cost 16 cores ~6 ms).  1 = always GPU. */
static int b200_poa_remsa_min_objects = 1;

/* called with the lock held once every active object waits */
static inline void b200_remsa_launch(b200_remsa_rv_t *rv){
	u4i j, n = rv->waiting;
	if(rv->in_bytes > rv->cap_in){   /* jobs that found no room (late) are copied here; the others' bytes move with realloc */
		rv->cap_in = rv->in_bytes + rv->in_bytes / 4 + 4096;
		rv->in = (uint8_t*)realloc(rv->in, rv->cap_in);
	}
	if(rv->match_ints > rv->cap_match){ rv->cap_match = rv->match_ints + rv->match_ints / 4 + 1024; free(rv->match); rv->match = (int32_t*)malloc(sizeof(int32_t) * rv->cap_match); }
	if(rv->in == NULL || rv->match == NULL){ fflush(stdout); fprintf(stderr, " -- Out of memory in %s -- %s:%d --\n", __FUNCTION__, __FILE__, __LINE__); fflush(stderr); abort(); }
	for(j=0;j<n;j++){
		b200_remsa_job_t *jb = rv->jobs[j];
		if(jb->late){
			memcpy(rv->in + jb->in_off, jb->src[0], 2 * (size_t)jb->sz1);
			memcpy(rv->in + jb->in_off + 2 * (size_t)jb->sz1, jb->src[1], 8 * (size_t)jb->sz1);
		}
		memcpy(rv->hdr + 8 * (size_t)j, jb->hdr, sizeof(jb->hdr));
		rv->in_off[j] = jb->in_off; rv->match_off[j] = jb->match_off;
	}
	if(B200_REMSA_BATCH_FN(rv->ctx, n, rv->hdr, rv->in, rv->in_off, rv->in_bytes, rv->match, rv->match_off, rv->match_ints, rv->out, NULL, NULL)){
		fflush(stdout); fprintf(stderr, " -- bsalign_b200: %s in %s -- %s:%d --\n", rv->ctx ? bsb200_last_error(rv->ctx) : "remsa batch failed", __FUNCTION__, __FILE__, __LINE__); fflush(stderr);
		abort();
	}
	for(j=0;j<n;j++) memcpy(rv->jobs[j]->out, rv->out + 4 * (size_t)j, sizeof(rv->jobs[j]->out));
	b200_poa_remsa_batches ++; b200_poa_remsa_jobs += n;
	rv->waiting = 0; rv->reserved = 0; rv->in_bytes = 0; rv->match_ints = 0; rv->gen ++;
	pthread_cond_broadcast(&rv->cv);
}

/* what B200_REMSA_CORE calls in an object thread: remsa_pedit_rd_bspoacore's contract (bspoa.h:3916-4045) - the read's nodes merged into
 * the msa nodes of the columns the walk matched them to, the walk's score returned - minus the DP matrices, which nothing reads afterwards */
static inline int b200_remsa_core_hook(void *gp, u2i rid, u4i rbeg, u4i rend, u1i **matrix, u1i **seqs, u1i *(*mats)[4], int mlen, int mbeg, int mend, int W){
	BSPOA *g = (BSPOA*)gp;
	b200_remsa_rv_t *rv = b200_remsa_rv;
	b200_remsa_job_t job;
	const u4i bw = W * WORDSIZE, HW = bw / 2;
	const int32_t *m;
	u4i roff;
	UNUSED(rbeg); UNUSED(matrix);
	memset(&job, 0, sizeof(job));
	job.hdr[0] = mlen; job.hdr[1] = bw; job.hdr[2] = mbeg; job.hdr[3] = mend; job.hdr[4] = rend;
	job.sz1 = roundup_times(mlen + bw, WORDSIZE);
	job.src[0] = seqs[0] - HW;       /* seqs[0], seqs[1]: bspoa.h:4209-4210, :4228-4229 */
	job.src[1] = mats[0][0] - HW;    /* mats[0][0..3], mats[1][0..3]: bspoa.h:4216-4223 */
	pthread_mutex_lock(&rv->mu);
	job.in_off = rv->in_bytes; job.match_off = rv->match_ints;
	rv->in_bytes += 10 * (uint64_t)job.sz1; rv->match_ints += rend;
	job.late = rv->in_bytes > rv->cap_in;
	rv->jobs[rv->reserved ++] = &job;
	pthread_mutex_unlock(&rv->mu);
	if(!job.late){   /* the arena only moves inside b200_remsa_launch, which needs this thread to wait first */
		memcpy(rv->in + job.in_off, job.src[0], 2 * (size_t)job.sz1);
		memcpy(rv->in + job.in_off + 2 * (size_t)job.sz1, job.src[1], 8 * (size_t)job.sz1);
	}
	pthread_mutex_lock(&rv->mu);
	rv->waiting ++;
	if(rv->waiting == rv->active) b200_remsa_launch(rv);
	else { unsigned long gen0 = rv->gen; while(rv->gen == gen0) pthread_cond_wait(&rv->cv, &rv->mu); }
	pthread_mutex_unlock(&rv->mu);
	if(job.out[1]){
		fflush(stdout); fprintf(stderr, " -- bsalign_b200: re-alignment of read %u left its band (status %d) in %s -- %s:%d --\n", (unsigned)rid, job.out[1], __FUNCTION__, __FILE__, __LINE__); fflush(stderr);
		abort();   /* the reference aborts here too (bspoa.h:3982-3985, :4027-4030) */
	}
	m = rv->match + job.match_off;
	for(roff=rend;roff-->0;){
		if(m[roff] >= 0){
			bspoanode_t *v = get_rdnode_bspoa(g, rid, roff);
			bspoanode_t *u = get_rdnode_bspoa(g, g->seqs->nseq + 1 + v->base, m[roff]);
			merge_nodes_bspoa(g, u, v);
		}
	}
	return job.out[0];
}

typedef struct { b200_poa_round_t *r; u4i k; } b200_remsa_thr_t;
static void *b200_remsa_object_thread(void *arg){
	b200_remsa_thr_t *t = (b200_remsa_thr_t*)arg;
	b200_remsa_rv_t *rv = b200_remsa_rv;
	b200_remsa_core_fn = b200_remsa_core_hook;
	b200_poa_round_object(t->r, t->k);
	b200_remsa_core_fn = NULL;
	pthread_mutex_lock(&rv->mu);
	rv->active --;
	if(rv->waiting && rv->waiting == rv->active) b200_remsa_launch(rv);
	pthread_mutex_unlock(&rv->mu);
	return NULL;
}

/* phase 2 (bspoa.h:4764-4777) of all objects, up to B200_REMSA_MAX_INFLIGHT at a time, one host thread each */
static inline void b200_poa_realign_run(b200_poa_round_t *r){
	b200_remsa_rv_t rv;
	pthread_t *th = (pthread_t*)malloc(sizeof(pthread_t) * B200_REMSA_MAX_INFLIGHT);
	b200_remsa_thr_t *ta = (b200_remsa_thr_t*)malloc(sizeof(b200_remsa_thr_t) * B200_REMSA_MAX_INFLIGHT);
	u4i beg, cnt, t;
	if((int)r->n < b200_poa_remsa_min_objects){ free(th); free(ta); b200_poa_round_run(r, 2); return; }
	memset(&rv, 0, sizeof(rv));
	pthread_mutex_init(&rv.mu, NULL); pthread_cond_init(&rv.cv, NULL);
	rv.ctx = r->ctx;
	rv.cap_jobs = B200_REMSA_MAX_INFLIGHT;
	rv.jobs = (b200_remsa_job_t**)malloc(sizeof(b200_remsa_job_t*) * rv.cap_jobs);
	rv.hdr = (int32_t*)malloc(sizeof(int32_t) * 12 * (size_t)rv.cap_jobs); rv.out = rv.hdr + 8 * (size_t)rv.cap_jobs;
	rv.in_off = (uint64_t*)malloc(sizeof(uint64_t) * 2 * (size_t)rv.cap_jobs); rv.match_off = rv.in_off + rv.cap_jobs;
	b200_remsa_rv = &rv;
	r->phase = 2; r->rid = 0;
	for(beg=0;beg<r->n;beg+=cnt){
		cnt = num_min(r->n - beg, (u4i)B200_REMSA_MAX_INFLIGHT);
		rv.active = cnt; rv.waiting = 0; rv.reserved = 0; rv.in_bytes = 0; rv.match_ints = 0;
		for(t=0;t<cnt;t++){
			ta[t].r = r; ta[t].k = beg + t;
			if(pthread_create(&th[t], NULL, b200_remsa_object_thread, ta + t)){
				fflush(stdout); fprintf(stderr, " -- cannot start a host thread per in-flight object in %s -- %s:%d --\n", __FUNCTION__, __FILE__, __LINE__); fflush(stderr);
				abort();
			}
		}
		for(t=0;t<cnt;t++) pthread_join(th[t], NULL);
	}
	b200_remsa_rv = NULL;
	free(rv.jobs); free(rv.hdr); free(rv.in_off); free(rv.in); free(rv.match); free(th); free(ta);
	pthread_mutex_destroy(&rv.mu); pthread_cond_destroy(&rv.cv);
}

/* the realn rounds + final msa / cns (bspoa.h:4764-4777) of n objects whose reads are all in their graphs */
static inline void b200_poa_realign_batch(bsb200_ctx *ctx, BSPOA **gs, u4i n){
	b200_poa_round_t rd;
	memset(&rd, 0, sizeof(rd));
	rd.ctx = ctx; rd.gs = gs; rd.n = n;
	b200_poa_realign_run(&rd);
}
#endif

static inline void b200_end_bspoa_batch(bsb200_ctx *ctx, BSPOA **gs, u4i n){
	b200_poa_pack_t *p = b200_poa_pack_init();
	u4i k, maxr = 0, *nheads, *ntails, *slot;
	u2i rid;
	int i;
	nheads = (u4i*)malloc(sizeof(u4i) * (n + 1)); ntails = (u4i*)malloc(sizeof(u4i) * (n + 1)); slot = (u4i*)malloc(sizeof(u4i) * (n + 1));
	{
		b200_poa_round_t rd;
		double lap[8] = {0, 0, 0, 0, 0, 0, 0, 0}, lap_t0, lap_t1;   /* where the wall time of a batch goes (printed with BSB200_HOSTPROF set) */
		struct timespec lap_ts;
#define B200_LAP(i) do { clock_gettime(CLOCK_MONOTONIC, &lap_ts); lap_t1 = lap_ts.tv_sec + 1e-9 * lap_ts.tv_nsec; lap[i] += lap_t1 - lap_t0; lap_t0 = lap_t1; } while(0)
		clock_gettime(CLOCK_MONOTONIC, &lap_ts); lap_t0 = lap_ts.tv_sec + 1e-9 * lap_ts.tv_nsec;
		rd.ctx = ctx; rd.gs = gs; rd.n = n; rd.nheads = nheads; rd.ntails = ntails; rd.slot = slot; rd.p = p;
		rd.mine = (b200_poa_pack_t**)calloc(n + 1, sizeof(b200_poa_pack_t*));
		rd.rid = 0;
		b200_poa_round_run(&rd, 4);   /* objects are independent here too */
		for(k=0;k<n;k++) if(gs[k]->seqs->nseq > 1 && gs[k]->nmsa > maxr) maxr = gs[k]->nmsa;
		B200_LAP(7);
#ifdef BSALIGN_B200_POA_KMER_H
		rd.kslot = (b200_poa_kmer_slot_t*)calloc(n + 1, sizeof(b200_poa_kmer_slot_t));
#endif
		for(rid=1;rid<maxr;rid++){   /* bspoa.h:4752-4763 with align_rd_bspoa (bspoa.h:2620-2667) split around the sweep */
			b200_poa_pack_clear(p);
			rd.rid = rid;
			b200_poa_round_run(&rd, 0); B200_LAP(0);
#ifdef BSALIGN_B200_POA_KMER_H
			b200_poa_kmer_round(&rd); B200_LAP(1);
			b200_poa_round_run(&rd, 3); B200_LAP(2);
#endif
			for(k=0;k<n;k++){   /* the batch is packed in object order */
				if(slot[k] != MAX_U4 - 2) continue;
				slot[k] = p->njobs;
				b200_poa_pack_append(p, rd.mine[k]);
			}
			B200_LAP(3);
#ifdef BSALIGN_B200_POA_HOST_TRACEBACK
			b200_poa_pack_run(ctx, p);
#else
			b200_poa_pack_run_walk(ctx, p);
#endif
			B200_LAP(4);
			b200_poa_round_run(&rd, 1); B200_LAP(5);
		}
		rd.rid = 0;
#ifdef BSALIGN_B200_POA_REMSA_H
		b200_poa_realign_run(&rd);   /* the re-alignment DPs of all objects as GPU batches (bsalign_b200_poa_remsa.h) */
#else
		b200_poa_round_run(&rd, 2);
#endif
		B200_LAP(6);
		if(getenv("BSB200_HOSTPROF")) fprintf(stderr, "[end_bspoa_batch] %u objects, %u rounds: reads into the graphs %.3f s, before the sweep %.3f s, k-mer batches %.3f s, band placement %.3f s, packing %.3f s, "
			"sweep + walk (GPU) %.3f s, graph surgery %.3f s, realn rounds + final msa / cns %.3f s\n", n, maxr, lap[7], lap[0], lap[1], lap[2], lap[3], lap[4], lap[5], lap[6]);
#ifdef BSALIGN_B200_POA_KMER_H
		free(rd.kslot);
#endif
		for(k=0;k<n;k++) if(rd.mine[k]) b200_poa_pack_free(rd.mine[k]);
		free(rd.mine);
	}
	free(nheads); free(ntails); free(slot);
	b200_poa_pack_free(p);
}

/* dump_binary_msa_bspoa (bspoa.h:1555-1584) through the library's writer: the columns are gathered in MSA order */
static inline void b200_dump_binary_msa_bspoa(BSPOA *g, char *metadat, u4i metalen, FILE *out){
	u4i nseq = g->nrds, mlen = g->msaidxs->size, mrow = g->seqs->nseq + 3, i;
	u1i *cols = (u1i*)malloc((size_t)mlen * (nseq + 1) + 1), *qa = (u1i*)malloc(2 * (size_t)mlen + 1);
	for(i=0;i<mlen;i++){
		u1i *col = g->msacols->buffer + g->msaidxs->buffer[i] * mrow;
		memcpy(cols + (size_t)i * (nseq + 1), col, nseq + 1);
		qa[i] = col[nseq + 1]; qa[mlen + i] = col[nseq + 2];
	}
	bsb200_msa_write(out, nseq, mlen, cols, qa, qa + mlen, metadat, metalen);
	free(cols); free(qa);
}

#ifdef BSALIGN_B200_OVERRIDE
static inline bsb200_ctx *bsalign_b200_poa_default_ctx(void){
	bsb200_ctx *ctx = bsb200_default_context();
	if(ctx == NULL){ fflush(stdout); fprintf(stderr, " -- bsalign_b200: no CUDA device, and there is no CPU fallback in %s -- %s:%d --\n", __FUNCTION__, __FILE__, __LINE__); fflush(stderr); abort(); }
	return ctx;
}
static inline void b200_end_bspoa(BSPOA *g){ b200_end_bspoa_batch(bsalign_b200_poa_default_ctx(), &g, 1); }
#define end_bspoa b200_end_bspoa
#endif

#endif
