// poa_backtrace.cuh -- device-side walk of alignment2graph_bspoa (bspoa.h:2274-2497).
//
// After the sweep, the reference walks back from (maxidx, maxoff) to the head, re-deriving every step from the node rows, and
// interleaves graph surgery (merge_nodes_bspoa) with the walk.  The surgery never changes what the walk reads (rows, reverse edge
// lists and their coverage, band offsets), so the walk can run on the device right behind the sweep and hand the host only its
// DECISIONS: for every read position the node it is matched to (or -1), plus the insertion / deletion counts and the end state.
// The host then replays merge_nodes_bspoa / cpos bookkeeping in the reference's order (include/bsalign_b200_poa_compat.h), and the
// 15 GB of row blocks per 1000-job round no longer cross PCIe.
//
// One thread per job: the walk is a chain of dependent lookups (reverse edges of the current node -> their rows), latency bound.
// Reverse-edge records carry the predecessor's band offset and base so that one lookup round reaches the row data.
#pragma once
#include "poa_kernels.cuh"

namespace bsb200 {

constexpr int kScoreMax = 536870911;   // SEQALIGN_SCORE_MAX, bsalign.h:59

struct PoaBtArgs {
	uint32_t njobs;
	const int32_t *par;
	const uint8_t *queries; const uint64_t *qoff; const uint32_t *slen;
	const uint64_t *node_off; const int2 *node;
	const int32_t *reoff;          // per job nnode + 1 entries at node_off[job] + job
	const uint64_t *redge_off;     // njobs + 1
	const int4 *rev;               // per reverse edge TWO records: {predecessor (local id), coverage, its rpos, base | bonus << 8} and
	                               // {the predecessor's own reverse-edge range begin, end, 0, 0}: one lookup round per step of the walk
	const uint32_t *head, *tail;
	const uint8_t *rows; const uint64_t *row_off;
	const int32_t *best;           // from the sweep: maxscr, maxidx, maxoff
	int32_t *match;                // per read position (arena indexed like queries)
	int32_t *trace;                // njobs x 8: final x, final node, mat, mis, ins, del, start node, flags
};

// reverse-edge records: predecessor id + coverage + the predecessor's node fields
__global__ void poa_rev_prep_kernel(uint32_t njobs, const uint64_t *node_off, const int2 *node, const int32_t *reoff, const uint64_t *redge_off,
		const int32_t *resrc, const int32_t *recov, int4 *rev){
	for(uint32_t job=blockIdx.x;job<njobs;job+=gridDim.x){
		const uint64_t e0 = redge_off[job], e1 = redge_off[job + 1];
		const int2 *nd = node + node_off[job];
		for(uint64_t e=e0+threadIdx.x;e<e1;e+=blockDim.x){
			const int w = resrc[e];
			const int2 r = nd[w];
			rev[2 * e] = make_int4(w, recov[e], r.x, (int)((((uint32_t)r.y >> 16) & 0xffu) | ((((uint32_t)r.y >> 24) & 1u) << 8)));
			rev[2 * e + 1] = make_int4(reoff[node_off[job] + job + w], reoff[node_off[job] + job + w + 1], 0, 0);
		}
	}
}

template<bool W8>   // W8: band of 128 cells (8 steps per lane), the reference's default: straight-line lookups with every load in flight at once
__global__ void __launch_bounds__(32) poa_backtrace_kernel(const PoaBtArgs a){
	// one job per warp, one walking lane: the walk is a chain of dependent lookups, and 32 diverged walks in one warp would serialise
	// (measured: 432 ms for 1000 jobs with a thread per job); the other lanes only clear the match array
	const uint32_t job = blockIdx.x;
	if(job >= a.njobs) return;
	const int32_t *par = a.par + (size_t)job * 10;
	const int bw = W8 ? 128 : par[0], W = W8 ? 8 : bw / kLanes;
	const int mode = par[1] & 3, Mm = par[2], Xx = par[3], O = par[4], E = par[5], Q = par[6], P = par[7], refbonus = par[9];
	const int pw = epi8_piecewise((int8_t)O, (int8_t)E, (int8_t)Q, (int8_t)P, bw);
	const uint32_t mmblk = ((uint32_t)bw * (pw + 1) + 68 + 15) / 16 * 16;
	const uint32_t slen = a.slen[job];
	const uint8_t *query = a.queries + a.qoff[job];
	int32_t * __restrict__ match = a.match + a.qoff[job];
	const uint64_t n0 = a.node_off[job];
	const uint32_t nnode = (uint32_t)(a.node_off[job + 1] - n0);
	const int2 *node = a.node + n0;
	const int32_t *reoff = a.reoff + n0 + job;
	const int4 *rev = a.rev + 2 * a.redge_off[job];
	const int head = (int)a.head[job], tail = (int)a.tail[job];
	const uint8_t * __restrict__ rows = a.rows + a.row_off[job];
	const int midx = a.best[(size_t)job * 3 + 1], xe = a.best[(size_t)job * 3 + 2];
	int32_t *out = a.trace + (size_t)job * 8;
	for(uint32_t k=threadIdx.x;k<slen;k+=32) match[k] = -1;
	__syncwarp();
	if(threadIdx.x) return;
	int flags = 0;
	if(midx < 0 || (uint32_t)midx >= nnode){
		out[0] = xe; out[1] = midx; out[2] = out[3] = out[4] = out[5] = 0; out[6] = midx; out[7] = 1;
		return;
	}
	// striped cell / anchors of a block
	#define CELL(blk, arr, p) ((int)(int8_t)(blk)[(size_t)(arr) * bw + ((p) % W) * 16 + (p) / W])
	#define ANCH(blk) ((const int32_t*)((blk) + (size_t)bw * (pw + 1)))
	auto getscore = [&](const uint8_t *blk, int pos) -> int {   // bsalign.h:3187-3197
		if(W8){
			// branch-free, all nine loads in flight at once (a dependent add per load would serialise the memory round trips); an
			// out-of-band position (the reference reads outside the row) is clamped for the loads and flagged
			const bool bad = pos < 0 || pos >= bw;
			const int pc = min(max(pos, 0), bw - 1);
			const int jj = pc >> 3, ii = pc & 7;
			int s = ANCH(blk)[jj];
			int b[8];
			#pragma unroll
			for(int k=0;k<8;k++) b[k] = (int)(int8_t)blk[k * 16 + jj];
			#pragma unroll
			for(int k=0;k<8;k++) s += (k <= ii) ? b[k] : 0;
			if(bad) flags |= 1;
			return bad ? kScoreMin : s;
		}
		if(pos < 0 || pos >= bw){ flags |= 1; return kScoreMin; }
		const int jj = pos / W, ii = pos - jj * W;
		int s = ANCH(blk)[jj];
		for(int k=0;k<=ii;k++) s += (int)(int8_t)blk[k * 16 + jj];
		return s;
	};
	int n = midx, nidx = midx, x = xe, bt = -1, Hs0 = 0, Hs1, Hs2 = 0, mat = 0, mis = 0, ins = 0, del = 0;
	int rpos_n = node[n].x, base_n = (int)(((uint32_t)node[n].y >> 16) & 0xffu), bonus_n = (int)(((uint32_t)node[n].y >> 24) & 1u);
	int re0 = reoff[n], re1 = reoff[n + 1];                  // reverse-edge range of n
	int nx_rpos = rpos_n, nx_base = base_n, nx_bonus = bonus_n, nx_e0 = re0, nx_e1 = re1;   // the same for nidx (set when a match is chosen)
	const uint8_t *ublk = rows + (size_t)n * mmblk; int urpos = rpos_n;     // the row the reference's `us` pointer refers to
	Hs1 = getscore(ublk, x - rpos_n);
	long long guard = 0; const long long guard_max = 8ll * ((long long)slen + nnode) + 64;
	while(true){
		if(++guard > guard_max){ flags |= 2; break; }
		if(n == head || x < 0) break;
		if(bt == 2 || bt == 4){
			// inside a deletion: leave n for a predecessor that explains the score (bspoa.h:2308-2356)
			bool found = false;
			del++;
			const int e0 = re0, e1 = re1;
			for(int ei=e0;ei<e1;ei++){
				const int4 r = rev[2 * ei], r2 = rev[2 * ei + 1];
				const int w = r.x, rw = r.z;
				if(x < rw || x >= rw + bw) continue;
				const uint8_t *blk = rows + (size_t)w * mmblk;
				ublk = blk; urpos = rw;
				const int p = x - rw;
				Hs0 = getscore(blk, p);
				int q;
				if(bt == 2) q = pw ? CELL(blk, 1, p) : (int)(int8_t)(O + E);
				else q = CELL(blk, 2, p);
				if(Hs0 + q != Hs1) continue;
				n = w; rpos_n = rw; base_n = r.w & 0xff; bonus_n = (r.w >> 8) & 1; re0 = r2.x; re1 = r2.y;
				if(q == ((bt == 2) ? O + E : Q + P)){ bt = -1; Hs1 = Hs0; Hs2 = 0; }
				else { Hs1 -= (bt == 2) ? E : P; Hs2++; }
				found = true;
				break;
			}
			if(!found){ flags |= 2; break; }
			continue;
		} else if(bt == 1 || bt == 3){
			// insertion run (bspoa.h:2357-2392)
			ins++;
			int t = O + E * Hs2;
			if(pw == 2) t = max(t, Q + P * Hs2);
			x--;
			if(Hs0 + t == Hs1){ bt = -1; Hs1 = Hs0; Hs2 = 0; }
			else if(x >= 0){
				const int p = x - urpos;
				if(p < 0 || p >= bw){ flags |= 1; break; }
				Hs0 -= CELL(ublk, 0, p);
				Hs2++;
			}
			continue;
		} else if(bt == 0){
			// match / mismatch of read position x with node n (bspoa.h:2393-2410)
			match[x] = n;
			if(n != head && n != tail && (int)query[x] == base_n) mat++; else mis++;
			x--;
			n = nidx; rpos_n = nx_rpos; base_n = nx_base; bonus_n = nx_bonus; re0 = nx_e0; re1 = nx_e1;
			bt = -1;
		} else {
			// decide the next step from the predecessors of n (bspoa.h:2411-2496)
			int btc = 0, bi = 0, b_w = 0, b_h0 = 0, bti_low = 0xFF;
			int4 b_r = make_int4(0, 0, 0, 0), b_r2 = make_int4(0, 0, 0, 0);
			bool have = false;
			const int e0 = re0, e1 = re1;
			const int qx = (int)query[x] & 3;
			const int hp = ((uint32_t)x + 1 < slen && query[x] != query[x + 1]) ? 1 : 0;
			for(int ei=e0;ei<e1;ei++){
				const int4 r = rev[2 * ei], r2 = rev[2 * ei + 1];
				const int w = r.x, cov = r.y, rw = r.z, base_w = r.w & 0xff;
				if(x < rw || x > bw + rw) continue;
				const uint8_t *blk = rows + (size_t)w * mmblk;
				ublk = blk; urpos = rw;
				int ft = 0;
				const int pcl = min(x - rw, bw - 1);                     // x == bw + rw: the deletion scores are forbidden below, the loads stay in the row
				const int cu = CELL(blk, 0, pcl), ce = pw ? CELL(blk, 1, pcl) : E, cq = pw == 2 ? CELL(blk, 2, pcl) : 0;
				const int anch0 = ANCH(blk)[0];
				if(x == bw + rw){ Hs0 = getscore(blk, x - rw - 1); ft |= (1 << 2) | (1 << 4); }
				else if(x == rw){
					Hs0 = anch0;
					if(rw == 0 && (mode == 1 || w == head)) ft |= 1 << 15; else ft |= 1 << 0;
				} else Hs0 = getscore(blk, x - rw - 1);
				const int kprof = (base_w == base_n) * 2 + bonus_n;
				int s = (int)(int8_t)((qx == base_n) ? ((kprof & 1) ? Mm + refbonus : Mm) : Xx);
				if(kprof < 2) s += hp;                                   // the hpc profiles, bsalign.h:2204-2206
				if(ft & (1 << 15)) s -= anch0;
				int scr[3];
				scr[0] = (ft & (1 << 0)) ? kScoreMin : s;
				scr[1] = (ft & (1 << 2)) ? kScoreMin : cu + ce;
				scr[2] = (ft & (1 << 4)) ? kScoreMin : (pw == 2 ? cu + cq : kScoreMax);
				#pragma unroll
				for(int i=0;i<3;i++){
					if(Hs0 + scr[i] == Hs1){
						if(cov > btc){ have = true; bi = i; bti_low = i; b_w = w; b_h0 = Hs0; btc = cov; b_r = r; b_r2 = r2; }
						else if(cov == btc && i == 0 && bti_low != 0){ have = true; bi = 0; bti_low = 0; b_w = w; b_h0 = Hs0; btc = cov; b_r = r; b_r2 = r2; }
					}
				}
			}
			if(!have){
				const int p = x - rpos_n;
				if(p < 0 || p >= bw){ flags |= 1; break; }
				bt = 1; Hs2 = 1;
				ublk = rows + (size_t)n * mmblk; urpos = rpos_n;
				Hs0 = Hs1 - CELL(ublk, 0, p);
			} else if(bi == 0){
				bt = 0; nidx = b_w; Hs1 = b_h0; Hs2 = 0;
				nx_rpos = b_r.z; nx_base = b_r.w & 0xff; nx_bonus = (b_r.w >> 8) & 1; nx_e0 = b_r2.x; nx_e1 = b_r2.y;
			}
			else if(bi == 1){ bt = 2; Hs2 = 1; }
			else { bt = 4; Hs2 = 1; }
		}
	}
	#undef CELL
	#undef ANCH
	out[0] = x; out[1] = n; out[2] = mat; out[3] = mis; out[4] = ins; out[5] = del; out[6] = midx; out[7] = flags;
}

} // namespace bsb200
