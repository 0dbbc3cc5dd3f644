// remsa_kernels.cuh -- re-alignment of reads against the MSA profile on the device: the DP and the walk of
// remsa_pedit_rd_bspoacore (bspoa.h:3916-4045; kernel maxmat_dp_diag_rowcal bspoa.h:3856-3896).
//
// One job = one read of one BSPOA object in MSA coordinates against the (reversed) consensus and the per-column profile scores, in the
// byte layout remsa_pedits_bspoa carves out of g->memp (bspoa.h:4209-4229): ten arrays of sz1 = roundup(mlen + bw, 16) bytes, each with
// bw / 2 bytes of padding in front: seqs0, seqs1, mats[0][0..3], mats[1][0..3].  The reference walks an anti-diagonal band of bw cells
// down the main diagonal (one SSE word per 16 cells) and keeps the DP in difference form in two byte matrices; the walk back re-derives
// every step from those differences.
//
// One WARP per job.  bw = 32 (the default: editbw / 2) is one cell per lane: the previous diagonal stays in registers, a cell's upper
// and left neighbours are one shuffle each, and the read / consensus / profile bytes of a cell slide along the lanes (one new byte per
// diagonal comes from memory, the other 31 from the neighbour lane).  Other widths keep two diagonals in shared memory.  Every diagonal
// is stored once (the walk needs it), the forward sweep leaves the step the
// walk takes from every cell as 2 bits (one 64-bit word per diagonal at bw = 32; the reference's two byte matrices are only written for
// checks), the walk is lane 0: a chain of dependent lookups in that small array.
#pragma once
#include "common.cuh"
#include <stdint.h>

namespace bsb200 {

struct RemsaArgs {
	uint32_t njobs;
	const int32_t *hdr;        // per job 8 ints: mlen, bw, mbeg, mend, rend (read positions), 0, 0, 0
	const uint8_t *in;         // input blocks
	const uint64_t *in_off;    // per job byte offset of its block
	uint8_t *mat;              // FULL only: per job two matrices of (2 * mlen + 1) * (bw + 2) bytes, back to back
	const uint64_t *mat_off;
	uint4 *codes;              // per job (2 * mlen + 1) rows of ceil(bw / 32) records: four bits per cell (see remsa_kernel)
	const uint64_t *code_off;  // in records
	int32_t *xs;               // scratch, laid out like match: the MSA position of every matched read position
	int32_t *match;            // per job rend ints: the MSA column a read position is matched to, or -1
	const uint64_t *match_off;
	int32_t *out;              // per job 4 ints: score of the walk, status, matched positions, 0
};

__device__ __forceinline__ int remsa_score(const uint8_t *seqs0, const uint8_t *seqs1, const uint8_t *mats0, const uint8_t *mats1, uint32_t sz1, int mlen, int xi, int yi){
	const int s1 = seqs1[mlen - 1 - yi], s0 = seqs0[xi];
	int h = (s1 < 4 ? mats0[(size_t)s1 * sz1 + xi] : 0) + (s0 < 4 ? mats1[(size_t)s0 * sz1 + (mlen - 1 - yi)] : 0);
	return h > 255 ? 255 : h;
}

constexpr int kRemsaWarps = 4;          // jobs per CTA
constexpr int kRemsaMaxBw = 256;        // widest band of the shared-memory path

// FULL: also store the two difference matrices of the reference (checks); the walk itself reads the bits the forward sweep leaves per
// cell, among them the step: 1 = the read position has no partner (x - 1), 2 = the column has none (y - 1), 3 = matched (x - 1, y - 1).  With
// s = H(cell) the reference's tests are s == f (H came from the left), s == e (from above), s == h (the cell's own score), in that order
// (bspoa.h:3998-4030); H = max(h, u, v), so one of them always holds.
template<bool FULL>
__global__ void __launch_bounds__(kRemsaWarps * 32) remsa_kernel(const RemsaArgs a){
	__shared__ uint8_t rowbuf[kRemsaWarps][2][2][kRemsaMaxBw + 2];   // [warp][matrix][parity][border + cells + border]
	const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t job = blockIdx.x * kRemsaWarps + wid;
	if(job >= a.njobs) return;
	const int32_t *hd = a.hdr + (size_t)job * 8;
	const int mlen = hd[0], bw = hd[1], mbeg = hd[2], mend = hd[3], rend = hd[4];
	const int half = bw / 2, rowlen = bw + 2;
	const uint32_t sz1 = (uint32_t)((mlen + bw + 15) / 16 * 16);
	const uint8_t *blk = a.in + a.in_off[job];
	const uint8_t *seqs0 = blk + half, *seqs1 = blk + sz1 + half, *mats0 = blk + 2 * (size_t)sz1 + half, *mats1 = blk + 6 * (size_t)sz1 + half;
	uint8_t *M0 = FULL ? a.mat + a.mat_off[job] : nullptr, *M1 = FULL ? M0 + (size_t)(2 * mlen + 1) * rowlen : nullptr;
	const int CW = (bw + 31) / 32;
	uint4 *codes = a.codes + a.code_off[job];
	int32_t *xs = a.xs + a.match_off[job];
	int32_t *match = a.match + a.match_off[job];
	int32_t *out = a.out + (size_t)job * 4;
	for(int c=lane;c<rend;c+=32) match[c] = -1;
	if(bw > kRemsaMaxBw || bw < 2 || (bw & 1) || mend <= mbeg || mbeg < 0 || mend > mlen){ if(lane == 0){ out[0] = 0; out[1] = 1; out[2] = 0; out[3] = 0; } return; }
	// ---- forward: diagonals 2 * mbeg .. 2 * mend - 2, row i + 1 of the matrices from row i (bspoa.h:3749-3758, 3856-3896, 3925-3934) ----
	if(FULL){
		uint8_t *r0 = M0 + (size_t)rowlen * 2 * mbeg, *r1 = M1 + (size_t)rowlen * 2 * mbeg;
		for(int c=lane;c<rowlen;c+=32){ r0[c] = (c == 1 + half - 1) ? 255 : 0; r1[c] = (c == 1 + half) ? 255 : 0; }
	}
	if(bw == 32){
		// one cell per lane; u / v of the previous diagonal in registers (borders: the values the reference writes beside the row)
		int pu = (lane == (uint32_t)half - 1) ? 255 : 0, pv = (lane == (uint32_t)half) ? 255 : 0;
		int bl_u = 0, bl_v = 0, br_u = 0, br_v = 0;   // border bytes left (-1) and right (bw) of the previous diagonal
		int x = mbeg, y = mbeg;
		for(int i=2*mbeg;;i++){
			const int dir = i & 1;
			const int h0 = remsa_score(seqs0, seqs1, mats0, mats1, sz1, mlen, x - half + (int)lane, y + half - (int)lane);
			int u, v;
			if(dir){ u = __shfl_down_sync(0xffffffffu, pu, 1); if(lane == 31) u = br_u; v = pv; }
			else { u = pu; v = __shfl_up_sync(0xffffffffu, pv, 1); if(lane == 0) v = bl_v; }
			int h = h0 < u ? u : h0;
			if(h < v) h = v;
			pu = h - v; pv = h - u;
			{
				const int t = (h == v && !(lane == 0 && dir == 0)) ? 1 : (h == u ? 2 : 3);
				const uint32_t lo = __ballot_sync(0xffffffffu, t & 1), hi = __ballot_sync(0xffffffffu, t & 2);
				const uint32_t nz = __ballot_sync(0xffffffffu, seqs0[x - half + (int)lane] < 4), hn = __ballot_sync(0xffffffffu, h0 != 0);
				if(lane == 0) codes[(size_t)(i + 1)] = make_uint4(lo, hi, nz, hn);
			}
			if(dir){ bl_u = 255; bl_v = 0; br_u = 0; br_v = 0; } else { bl_u = 0; bl_v = 0; br_u = 0; br_v = 255; }
			if(FULL){
				uint8_t *nu = M0 + (size_t)rowlen * (i + 1) + 1, *nv = M1 + (size_t)rowlen * (i + 1) + 1;
				nu[lane] = (uint8_t)pu; nv[lane] = (uint8_t)pv;
				if(lane == 0){ nu[-1] = (uint8_t)bl_u; nv[-1] = (uint8_t)bl_v; nu[bw] = (uint8_t)br_u; nv[bw] = (uint8_t)br_v; }
			}
			if(dir) y++; else x++;
			if(x >= mend) break;
		}
		(void)br_v; (void)bl_u;
	} else {
		uint8_t (*rb)[2][kRemsaMaxBw + 2] = rowbuf[wid];
		for(int c=lane;c<rowlen;c+=32){ rb[0][0][c] = (c == 1 + half - 1) ? 255 : 0; rb[1][0][c] = (c == 1 + half) ? 255 : 0; }
		__syncwarp();
		int x = mbeg, y = mbeg, par = 0;
		for(int i=2*mbeg;;i++,par^=1){
			const int dir = i & 1;
			const uint8_t *pu = rb[0][par] + 1, *pv = rb[1][par] + 1;
			uint8_t *qu = rb[0][par ^ 1] + 1, *qv = rb[1][par ^ 1] + 1;
			uint8_t *nu = FULL ? M0 + (size_t)rowlen * (i + 1) + 1 : nullptr, *nv = FULL ? M1 + (size_t)rowlen * (i + 1) + 1 : nullptr;
			for(int c0=0;c0<bw;c0+=32){
				const int c = c0 + (int)lane;
				int t = 0; bool cnz = false, chn = false;
				if(c < bw){
					int h = remsa_score(seqs0, seqs1, mats0, mats1, sz1, mlen, x - half + c, y + half - c);
					cnz = seqs0[x - half + c] < 4; chn = h != 0;
					const int u = dir ? pu[c + 1] : pu[c], v = dir ? pv[c] : pv[c - 1];
					if(h < u) h = u;
					if(h < v) h = v;
					qu[c] = (uint8_t)(h - v); qv[c] = (uint8_t)(h - u);
					if(FULL){ nu[c] = (uint8_t)(h - v); nv[c] = (uint8_t)(h - u); }
					t = (h == v && !(c == 0 && dir == 0)) ? 1 : (h == u ? 2 : 3);
				}
				const uint32_t lo = __ballot_sync(0xffffffffu, t & 1), hi = __ballot_sync(0xffffffffu, t & 2);
				const uint32_t nz = __ballot_sync(0xffffffffu, cnz), hn = __ballot_sync(0xffffffffu, chn);
				if(lane == 0) codes[(size_t)(i + 1) * CW + (c0 >> 5)] = make_uint4(lo, hi, nz, hn);
			}
			if(lane == 0){
				const uint8_t lu = dir ? 255 : 0, rv = dir ? 0 : 255;
				qu[-1] = lu; qv[-1] = 0; qu[bw] = 0; qv[bw] = rv;
				if(FULL){ nu[-1] = lu; nv[-1] = 0; nu[bw] = 0; nv[bw] = rv; }
			}
			__syncwarp();
			if(dir) y++; else x++;
			if(x >= mend) break;
		}
	}
	__threadfence_block();
	__syncwarp();
	// ---- the walk (bspoa.h:3962-4040): lane 0, on the four bits the sweep left per cell: step (2 bits: 1 = x - 1, 2 = y - 1, 3 = both),
	// "the read has a base in this column", "the cell's own score is not zero".  The score of the walk is the sum of the matched cells'
	// scores: those of matched read positions are added up by the whole warp afterwards, the others (no read base there: zero on the
	// reference's own inputs) on the spot.
	int scr = 0, err = 0, nmatch = 0;
	if(lane == 0 && !(hd[5] & 1)){   // (hdr[5] bit 0: skip the walk - a development aid for timing the sweep alone)
		int xi = mend - 1, yi = mend - 1, roff = rend;
		while(xi >= 0 && yi >= 0){
			const int i = xi + yi;
			if(i < mbeg + mbeg) break;
			const int dir = i & 1;
			const int xx = (xi - yi - dir) / 2 + half;
			if(xx < 0 || xx >= bw){ err |= 1; break; }
			const uint4 wc = codes[(size_t)(i + 1) * CW + (xx >> 5)];
			// the walk only moves down the diagonals: ask for the records it reaches ~48 diagonals from now
			if(i > mbeg + mbeg + 48) asm volatile("prefetch.global.L1 [%0];" :: "l"(codes + (size_t)(i + 1 - 48) * CW + (xx >> 5)));
			const uint32_t bit = 1u << (xx & 31);
			const int t = ((wc.x & bit) ? 1 : 0) | ((wc.y & bit) ? 2 : 0);
			const bool nz = (wc.z & bit) != 0;
			if(t == 1){ if(nz) roff--; xi--; }
			else if(t == 2){ yi--; }
			else if(t == 3){
				if(nz){ roff--; if(roff >= 0 && roff < rend){ match[roff] = yi; xs[roff] = xi; nmatch++; } else err |= 1; }
				else if(wc.w & bit) scr += remsa_score(seqs0, seqs1, mats0, mats1, sz1, mlen, xi, yi);
				xi--; yi--;
			} else { err |= 2; break; }
		}
	}
	__threadfence_block();
	__syncwarp();
	for(int r=lane;r<rend;r+=32){ const int yj = match[r]; if(yj >= 0) scr += remsa_score(seqs0, seqs1, mats0, mats1, sz1, mlen, xs[r], yj); }
	for(int o=16;o;o>>=1) scr += __shfl_xor_sync(0xffffffffu, scr, o);
	if(lane == 0){ out[0] = scr; out[1] = err; out[2] = nmatch; out[3] = 0; }
}

} // namespace bsb200
