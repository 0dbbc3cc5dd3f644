// common.cuh -- constants and small device helpers shared by the bsalign_b200 kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bsb200 {

constexpr int kGroup = 8;             // threads per pair in the epi8 forward kernel
constexpr int kLanes = 16;            // SSE lanes of the reference build (WORDSIZE, bsalign.h:142)
constexpr int kMetaInts = 20;         // per-row anchors: ub[17], rbeg, 2 pad  (80 B, bsalign.h:3878)
constexpr int kScoreMin = -536870911; // SEQALIGN_SCORE_MIN, bsalign.h:58
constexpr int kEpi8Min = -63;         // SEQALIGN_SCORE_EPI8_MIN, bsalign.h:56
constexpr int kEpi8Max = 63;
// internal per-pair status bits between the forward and the traceback kernel (never returned to the caller)
constexpr int kStSkew = 0x10000000;   // the pair's trace is in the skewed layout of the wavefront kernel (epi8_wave.cuh)
constexpr int kStRedo = 0x20000000;   // a guard of the wavefront kernel tripped: the two-pass kernel redoes the pair

// ---- s16x2 helpers: two int8 lanes per register ------------------------------------------------------
__device__ __forceinline__ uint32_t pk(int lo, int hi){ return (uint32_t)(lo & 0xffff) | ((uint32_t)hi << 16); }
__device__ __forceinline__ uint32_t pk1(int v){ return pk(v, v); }
// raw PRMT: selector nibble 8|i replicates the sign of byte i (__byte_perm() masks that bit away)
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel){ uint32_t d; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel)); return d; }
__device__ __forceinline__ int lo16(uint32_t v){ return (int)(short)(v & 0xffff); }
__device__ __forceinline__ int hi16(uint32_t v){ return (int)(short)(v >> 16); }
__device__ __forceinline__ int clamp8(int v){ return max(-128, min(127, v)); }

__host__ __device__ __forceinline__ int epi8_piecewise(int go1, int ge1, int go2, int ge2, int bw){
	if(go2 < go1 && ge2 > ge1 && go2 + ge2 < go1 + ge1 && (go1 - go2) / (ge1 - ge2) < bw) return 2; // bsalign.h:2084-2092
	return go1 ? 1 : 0;
}

// Row image layout shared by the forward kernel, the HBM trace and the traceback kernel.  One array (u, e or q)
// of a row is ceil(W/8) chunks of 128 bytes; chunk c holds, for each of the 8 threads t of the group, 16 bytes =
// the byte pairs (lane 2t, lane 2t+1) of steps 8c..8c+7.  A group's 128-bit access to one chunk therefore covers
// all 32 shared-memory banks exactly once, and the image is what is streamed to HBM unchanged.
__host__ __device__ __forceinline__ uint32_t epi8_image_bytes(uint32_t W){ return (W + 7) / 8 * 128; }
// byte offset of band position p (lane j = p / W, step i = p % W) inside one array image
__host__ __device__ __forceinline__ uint32_t epi8_cell_offset(uint32_t j, uint32_t i){ return (i >> 3) * 128 + (j >> 1) * 16 + (i & 7) * 2 + (j & 1); }

// The wavefront kernel (epi8_wave.cuh) keeps the two steps of a word step-major per lane - bytes (A_k, A_k+1, B_k, B_k+1) instead
// of (A_k, B_k, A_k+1, B_k+1) - so that two s16x2 results with values in 0..255 interleave into a word with one IMAD (lo + 256 * hi),
// and stores e biased by +128 like u.  Traces in the skewed layout use this byte order.
__host__ __device__ __forceinline__ uint32_t epi8_cell_offset_w(uint32_t j, uint32_t i){ return (i >> 3) * 128 + (j >> 1) * 16 + ((i & 7) >> 1) * 4 + (j & 1) * 2 + (i & 1); }
// ... and in the HBM trace of those pairs a thread's 16 bytes of u and its 16 bytes of e of a chunk sit side by side (256 bytes per chunk):
// byte offset of u of (lane j, step i) inside a slot; e is 16 bytes further
__host__ __device__ __forceinline__ uint32_t epi8_trace_offset_w(uint32_t j, uint32_t i){ return (i >> 3) * 256 + (j >> 1) * 32 + ((i & 7) >> 1) * 4 + (j & 1) * 2 + (i & 1); }

// Sub-lane anchors for the traceback: besides the 17 block anchors of the reference, a row carries the absolute score
// at the end of every 32nd step of every lane (int32 [g-1][lane], g = 1 .. ngrp-1), so that a score lookup sums at
// most 32 cells (4 chunks = 4 sectors) instead of a whole lane.
constexpr uint32_t kAnchorSteps = 32, kAnchorChunks = kAnchorSteps / 8;
constexpr uint32_t kStageAlign = 32;   // sub-blocks of the wavefront kernel are whole 32-step groups (the chunks of row_max, bsalign.h:3227)
__host__ __device__ __forceinline__ uint32_t epi8_anchor_groups(uint32_t W){ return (W + kAnchorSteps - 1) / kAnchorSteps; }
__host__ __device__ __forceinline__ uint32_t epi8_anchor_bytes(uint32_t W){ return ((epi8_anchor_groups(W) - 1) * 64 + 127) / 128 * 128; }   // rows stay 128-byte aligned
// The two-pass kernel computes the anchors in a loop of its own behind pass 2: they pay off only when the widest lane of a batch
// exceeds 64 steps.  The wavefront kernel has the running sums in registers anyway and writes them whenever a lane has more than
// one anchor group (written between the chunk groups, rows padded to whole 128-byte lines).  Config 2 (W = 63, one anchor per lane
// and row): forward 52.4 -> 53.9 ms, walk 16.9 -> 12.1 ms.
__host__ __device__ __forceinline__ bool epi8_use_anchors(uint32_t maxW){ return maxW > 64; }
__host__ __device__ __forceinline__ bool epi8_wave_use_anchors(uint32_t maxW){ return maxW > kAnchorSteps; }
__host__ __device__ __forceinline__ uint32_t epi8_row_bytes(uint32_t W, int pw){ return epi8_image_bytes(W) * (pw + 1) + epi8_anchor_bytes(W); }

struct CigarSink {
	uint32_t *buf; uint32_t cap, n, run; int err;
	__device__ __forceinline__ void put(uint32_t w){ if(buf){ if(n < cap) buf[n] = w; else err |= 4; } n++; }
	__device__ __forceinline__ void push(uint32_t op, uint32_t sz){ // bsalign.h:409-417
		if(op == (run & 0xf)){ run += sz << 4; return; }
		if(run) put(run);
		run = sz << 4 | op;
	}
	__device__ __forceinline__ void flush(){ if(run) put(run); run = 0; }
};

// The walk emits operations end-to-start; the final cigar is their reverse (bsalign.h:3850, :1042).  Each pair
// reserves its exact length in one dense arena so that only real cigar words travel back over PCIe.
__device__ __forceinline__ void emit_dense(const CigarSink &cg, uint32_t *dense, uint64_t *dense_off, unsigned long long *dense_total, uint32_t *ncigar, uint32_t pair){
	// the count that lays out the dense arena is the clamped one (a walk only outgrows its qlen + tlen + 2 scratch on pairs that
	// are flagged anyway; cg.err then carries BSB200_ST_CIGCAP)
	uint32_t n = (cg.buf && cg.n > cg.cap) ? cg.cap : cg.n;
	if(ncigar) ncigar[pair] = n;
	if(!cg.buf || !dense) return;
	unsigned long long off = atomicAdd(dense_total, (unsigned long long)n);
	dense_off[pair] = off;
	for(uint32_t i=0;i<n;i++) dense[off + i] = cg.buf[n - 1 - i];
}

} // namespace bsb200
