/*
 * bsalign_b200_poa_remsa.h -- lets b200_end_bspoa_batch (bsalign_b200_poa_compat.h) run the read re-alignment DP of remsa_pedits_bspoa
 * (bspoa.h:4178-4457: per read one remsa_pedit_rd_bspoacore, bspoa.h:3916-4045) for all in-flight objects as GPU batches
 * (bsb200_remsa_batch: one job per object and rendezvous).
 *
 * remsa_pedit_rd_bspoacore is defined AND called inside bspoa.h, so unlike the k-mer hook (bsalign_b200_poa_kmer.h) a macro from outside
 * cannot redirect the call: the reference-side change is ONE line, the callee's name at bspoa.h:4451,
 *
 *     -		scr = remsa_pedit_rd_bspoacore(g, rid, qb, qe, matrix, seqs, mats, mlen, mbeg, mend, W);
 *     +		scr = B200_REMSA_CORE(g, rid, qb, qe, matrix, seqs, mats, mlen, mbeg, mend, W);
 *
 * with this header included BETWEEN the reference's two headers (it only needs bsalign.h's integer types):
 *     #include "bsalign.h"
 *     #include "bsalign_b200_poa_remsa.h"
 *     #include "bspoa.h"
 *     #include "bsalign_b200_poa_compat.h"
 * (a tree without this header keeps compiling with  #ifndef B200_REMSA_CORE / #define B200_REMSA_CORE remsa_pedit_rd_bspoacore / #endif
 * in front of remsa_pedits_bspoa).  B200_REMSA_CORE calls the reference's own function unless the calling THREAD has a hook installed,
 * which only the object threads of b200_end_bspoa_batch do: end_bspoa of the reference itself and every other caller behave as before.
 * oracle/Makefile (target dropin_remsa) applies exactly that one-line change to a temporary copy of bspoa.h when it builds the test
 * program; nothing of the reference is kept in this repository.
 */
#ifndef BSALIGN_B200_POA_REMSA_H
#define BSALIGN_B200_POA_REMSA_H

#include <stdint.h>

/* the argument list of remsa_pedit_rd_bspoacore (bspoa.h:3916) with the object as void* (BSPOA is not declared yet) and the arrays decayed */
typedef int (*b200_remsa_core_fn_t)(void *g, u2i rid, u4i rbeg, u4i rend, u1i **matrix, u1i **seqs, u1i *(*mats)[4], int mlen, int mbeg, int mend, int W);
static __thread b200_remsa_core_fn_t b200_remsa_core_fn = NULL;

#define B200_REMSA_CORE(g, rid, rbeg, rend, matrix, seqs, mats, mlen, mbeg, mend, W) \
	(b200_remsa_core_fn ? b200_remsa_core_fn((g), (rid), (rbeg), (rend), (matrix), (seqs), (mats), (mlen), (mbeg), (mend), (W)) \
	                    : remsa_pedit_rd_bspoacore((g), (rid), (rbeg), (rend), (matrix), (seqs), (mats), (mlen), (mbeg), (mend), (W)))

#endif
