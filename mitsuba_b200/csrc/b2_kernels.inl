// Wavefront kernels.  This file is compiled twice: kernels_parity.cu (-fmad=false, namespace
// b2::parity) and kernels_fast.cu (FMA contraction on, namespace b2::fast).  B2_KNS names the namespace.
//
// One iteration of the host loop = one bounce stage for every in-flight path:
//   k_generate : splat finished paths into the film (ImageBlock::put, imageblock.h:124-204) and refill
//                their slots with new (pixel, sample) work: sampler->generate/advance, camera ray
//                (integrator.cpp:162-187, perspective.cpp:271-298)
//   k_extend   : closest-hit query of every live path (skdtree.cpp:112-142) [+ material-class binning]
//   k_shade<C> : the body of MIPathTracer::Li's loop (path.cpp:135-287) for one material class:
//                emitter-hit MIS + Russian roulette of the previous bounce, then emission, NEE sampling
//                (emits a shadow ray into the compacted shadow queue) and BSDF sampling (next ray)
//   k_occluded : any-hit query of the shadow queue (skdtree.cpp:207-226); unoccluded -> Li += contribution
#include "b2_math.cuh"
#include "b2_types.h"
#include "b2_sampler.cuh"
#include "b2_bsdf.cuh"
#include "b2_trace.cuh"
#include "b2_medium.cuh"
#include "b2_texture.cuh"
#include "b2_envmap.cuh"
#include "b2_launch.h"

namespace b2 {
namespace B2_KNS {

// ------------------------------------------------------------------------------------------------
// TMA staging of the head of the node / triangle arrays into shared memory (cp.async.bulk + mbarrier)
// ------------------------------------------------------------------------------------------------
B2_DEV uint32_t smemAddr(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

B2_DEV void stageScene(const DScene &sc, float4 *sNodes, float4 *sTris, uint64_t *bar, bool wide = false) {
    const uint32_t nb = wide ? sc.stageNodes8 * 80u : sc.stageNodes * 64u;
    const void *nodeSrc = wide ? (const void *) sc.nodes8 : (const void *) sc.nodes;
    const uint32_t barA = smemAddr(bar);
#ifdef B2_FAST_TRI
    // flat leaf: the paired records (never larger than the triangle list they replace); BVH: head of the plane array
    const float4 *triSrc = sc.rootCount ? sc.flatRec : sc.triPlane;
    const uint32_t tb = sc.rootCount ? sc.flatBytes : sc.stageTris * 48u;
#else
    const float4 *triSrc = sc.triAccel;
    const uint32_t tb = sc.stageTris * 48u;
#endif
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barA));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barA), "r"(nb + tb) : "memory");
        if (nb)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(sNodes)),
                         "l"(nodeSrc), "r"(nb), "r"(barA)
                         : "memory");
        if (tb)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(sTris)),
                         "l"(triSrc), "r"(tb), "r"(barA)
                         : "memory");
    }
    // every thread waits for phase 0
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(barA)
            : "memory");
    }
}

// dynamic shared memory carve-up for the traversal kernels
// (wide: the 8-ary tree -- uint2 stack entries, 80-byte nodes staged -- for the kernels that walk it: k_extend / k_occluded / k_trace_rays
// on non-instanced BVH scenes)
B2_DEV TraceMem setupTraceMem(const DScene &sc, unsigned char *smem, bool wide = false) {
    uint32_t *stack = (uint32_t *) smem;
    size_t off = wide ? (size_t) B2_STACK8_DEPTH * blockDim.x * sizeof(uint2) : (size_t) B2_STACK_DEPTH * blockDim.x * sizeof(uint32_t);
    off = (off + 127) & ~(size_t) 127;
    float4 *sNodes = (float4 *) (smem + off);
    off += wide ? (size_t) sc.stageNodes8 * 80 : (size_t) sc.stageNodes * 64;
    float4 *sTris = (float4 *) (smem + off);
    off += (size_t) sc.stageTriBytes;
    off = (off + 15) & ~(size_t) 15;
    uint64_t *bar = (uint64_t *) (smem + off);
    stageScene(sc, sNodes, sTris, bar, wide);
    TraceMem tm;
    tm.gNodes8 = (const float4 *) sc.nodes8;
    tm.sNodes8 = sNodes;
    tm.stageNodes8 = wide ? sc.stageNodes8 : 0u;
    tm.stack8 = (uint2 *) smem + threadIdx.x;
    tm.gNodes = (const float4 *) sc.nodes;
#ifdef B2_FAST_TRI
    tm.gTris = sc.triPlane;
#else
    tm.gTris = sc.triAccel;
#endif
    tm.sNodes = sNodes;
    tm.sTris = sTris;
    tm.stageNodes = sc.stageNodes;
    tm.stageTris = sc.stageTris;
    tm.stack = stack + threadIdx.x;
    tm.stride = blockDim.x;
    return tm;
}

// ------------------------------------------------------------------------------------------------
// cp.async (LDGSTS) staging of per-slot records: the thread that will work on a slot in the NEXT loop iteration issues the copies
// of that slot's pool records into its own shared-memory cells now and waits for them one iteration later, so the HBM latency of
// the scattered / streamed 16-byte records overlaps with the arithmetic of the current item instead of stalling the warp
// (ncu, round 1: k_shade and k_generate were long-scoreboard bound at 28-32 % of the DRAM peak).  Every thread reads back only
// what it copied itself, so cp.async.wait_group is the only synchronisation needed.
// ------------------------------------------------------------------------------------------------
B2_DEV void cpAsync16(void *smemDst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smemAddr(smemDst)), "l"(gsrc) : "memory");
}
B2_DEV void cpAsync8(void *smemDst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smemAddr(smemDst)), "l"(gsrc) : "memory");
}
B2_DEV void cpAsync4(void *smemDst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smemAddr(smemDst)), "l"(gsrc) : "memory");
}
B2_DEV void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> B2_DEV void cpAsyncWait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

B2_DEV unsigned long long globalTimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// per-launch duration = max over CTAs of the end stamp - min over CTAs of the start stamp
B2_DEV void stampBegin(const DRender &rp, uint32_t it, int stage) {
    if (rp.stampStart && threadIdx.x == 0 && it < B2_MAX_STAMPS) atomicMin(rp.stampStart + (size_t) it * 4 + stage, globalTimer());
}
B2_DEV void stampEnd(const DRender &rp, uint32_t it, int stage) {
    if (rp.stampEnd && threadIdx.x == 0 && it < B2_MAX_STAMPS) atomicMax(rp.stampEnd + (size_t) it * 4 + stage, globalTimer());
}

B2_DEV uint32_t warpSum(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// warp-ballot stream compaction: lanes with `want` get consecutive indices from *counter (one atomic per warp)
B2_DEV unsigned long long warpAppend64(bool want, unsigned long long *counter) {
    const unsigned active = __activemask();
    const unsigned mask = __ballot_sync(active, want);
    if (!want) return ~0ull;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned long long) __popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + (unsigned long long) __popc(mask & ((1u << lane) - 1u));
}
// 64-bit atomic add whose (low word of the) result is consumed later: nothing between the issue and the first use waits for it
B2_DEV uint32_t atomAddPending(unsigned long long *counter, uint32_t v) {
    unsigned long long r;
    asm volatile("atom.global.add.u64 %0, [%1], %2;" : "=l"(r) : "l"(counter), "l"((unsigned long long) v) : "memory");
    return (uint32_t) r;
}
B2_DEV uint32_t warpAppend(bool want, unsigned long long *counter) { return (uint32_t) warpAppend64(want, counter); }
// two independent appends whose atomics are issued back to back (their latencies overlap); all 32 lanes must call
B2_DEV void warpAppend2(bool wantA, unsigned long long *ctrA, bool wantB, unsigned long long *ctrB, uint32_t &idxA, uint32_t &idxB) {
    const unsigned mA = __ballot_sync(0xffffffffu, wantA), mB = __ballot_sync(0xffffffffu, wantB);
    const int lane = threadIdx.x & 31;
    unsigned long long bA = 0, bB = 0;
    if (lane == 0) {
        if (mA) bA = atomicAdd(ctrA, (unsigned long long) __popc(mA));
        if (mB) bB = atomicAdd(ctrB, (unsigned long long) __popc(mB));
    }
    bA = __shfl_sync(0xffffffffu, bA, 0);
    bB = __shfl_sync(0xffffffffu, bB, 0);
    idxA = (uint32_t) bA + __popc(mA & ((1u << lane) - 1u));
    idxB = (uint32_t) bB + __popc(mB & ((1u << lane) - 1u));
}

// ------------------------------------------------------------------------------------------------
// film: ImageBlock::put(pos, spec, alpha) with global atomics (imageblock.h:124-204, full-frame form)
// ------------------------------------------------------------------------------------------------
B2_DEV float evalDiscretized(const DFilter &f, float x) { // rfilter.h:76-77
    int i = (int) fabsf(x * f.scaleFactor);
    return f.values[i < 31 ? i : 31];
}
B2_DEV void redAddV4(float4 *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// The footprint arithmetic is relative to the 32x32 render block (scene.cpp:24) that owns the generating pixel,
// exactly like ImageBlock::put: pos = _pos - 0.5 - (offset - border).  (Full-frame arithmetic would round
// differently and occasionally pick a neighbouring filter bin.)  Block + border pixels outside the film are
// dropped, as Bitmap::accumulate clips them when the block is merged (imageblock.h:103-107).
B2_DEV bool filmPut(const DFilter &f, float4 *filmRGBA, float *filmW, int W, int H, float posX, float posY, int pixX, int pixY,
                    const V3 &spec, float alpha) {
    const float value[5] = {spec.x, spec.y, spec.z, alpha, 1.0f};
#pragma unroll
    for (int i = 0; i < 5; ++i)
        if (!isfinite(value[i]) || value[i] < 0) return false;
    const int border = f.borderSize;
    const int ox = (pixX >> 5) << 5, oy = (pixY >> 5) << 5;
    const int sizeX = min(32, W - ox) + 2 * border, sizeY = min(32, H - oy) + 2 * border;
    const float px = posX - 0.5f - (float) (ox - border), py = posY - 0.5f - (float) (oy - border);
    const int minx = max((int) ceilf(px - f.radius), 0), miny = max((int) ceilf(py - f.radius), 0);
    const int maxx = min((int) floorf(px + f.radius), sizeX - 1), maxy = min((int) floorf(py + f.radius), sizeY - 1);
    for (int y = miny; y <= maxy; ++y) {
        const int fy = oy - border + y;
        if (fy < 0 || fy >= H) continue;
        const float wy = evalDiscretized(f, (float) y - py);
        for (int x = minx; x <= maxx; ++x) {
            const int fx = ox - border + x;
            if (fx < 0 || fx >= W) continue;
            const float w = evalDiscretized(f, (float) x - px) * wy;
            const size_t o = (size_t) fy * W + fx;
            // one 16-byte reduction per pixel: (r, g, b, weight * alpha); the rest of the weight, w (1 - alpha), goes to a second plane and
            // is zero -- no second atomic -- whenever the camera ray hit something (k_film_pack: weight = both parts, no cancellation)
            redAddV4(filmRGBA + o, w * value[0], w * value[1], w * value[2], w * (value[4] * value[3]));
            if (value[3] != 1.0f) atomicAdd(filmW + o, w * (value[4] * (1.0f - value[3])));
        }
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// camera: perspective.cpp:271-298
// ------------------------------------------------------------------------------------------------
B2_DEV void cameraRay(const DCamera &cam, float sx, float sy, V3 &o, V3 &d, float &mint, float &maxt, float apx = 0.5f, float apy = 0.5f) {
    const float *M = cam.sampleToCamera;
    float px = sx * cam.invResX, py = sy * cam.invResY;
    float x = M[0] * px + M[1] * py + M[2] * 0.0f + M[3];
    float y = M[4] * px + M[5] * py + M[6] * 0.0f + M[7];
    float z = M[8] * px + M[9] * py + M[10] * 0.0f + M[11];
    float w = M[12] * px + M[13] * py + M[14] * 0.0f + M[15];
    V3 nearP(x, y, z);
    if (w != 1.0f) nearP = nearP / w;
    if (cam.apertureRadius > 0) { // thinlens.cpp:327-350: aperture point on the lens disk, direction through the point in focus
        float tx, ty;
        squareToUniformDiskConcentric(apx, apy, tx, ty);
        tx *= cam.apertureRadius; ty *= cam.apertureRadius;
        const V3 apertureP(tx, ty, 0.0f);
        const V3 focusP = nearP * (cam.focusDistance / nearP.z);
        const V3 dl = normalize(focusP - apertureP);
        const float invZ = 1.0f / dl.z;
        const float *T = cam.camToWorld;
        o = V3(T[0] * tx + T[1] * ty + T[2] * 0.0f + T[3], T[4] * tx + T[5] * ty + T[6] * 0.0f + T[7], T[8] * tx + T[9] * ty + T[10] * 0.0f + T[11]);
        d = V3(T[0] * dl.x + T[1] * dl.y + T[2] * dl.z, T[4] * dl.x + T[5] * dl.y + T[6] * dl.z, T[8] * dl.x + T[9] * dl.y + T[10] * dl.z);
        mint = cam.nearClip * invZ;
        maxt = cam.farClip * invZ;
        return;
    }
    V3 dl = normalize(nearP);
    float invZ = 1.0f / dl.z;
    const float *T = cam.camToWorld;
    d = V3(T[0] * dl.x + T[1] * dl.y + T[2] * dl.z, T[4] * dl.x + T[5] * dl.y + T[6] * dl.z, T[8] * dl.x + T[9] * dl.y + T[10] * dl.z);
    o = V3(cam.origin[0], cam.origin[1], cam.origin[2]);
    mint = cam.nearClip * invZ;
    maxt = cam.farClip * invZ;
}

// Sensor::sampleRayDifferential + RayDifferential::scaleDifferential(scale): the ray of cameraRay() plus the directions of the two
// offset rays through the neighbouring pixels (their origin is the ray origin for both sensors)
B2_DEV void cameraRayDifferential(const DCamera &cam, float spx, float spy, float apx, float apy, float scale, V3 &o, V3 &d, V3 &rxD, V3 &ryD) {
    float mint, maxt;
    cameraRay(cam, spx, spy, o, d, mint, maxt, apx, apy);
    const float *M = cam.sampleToCamera;
    const float px = spx * cam.invResX, py = spy * cam.invResY, pz = 0.0f;
    const float w = M[12] * px + M[13] * py + M[14] * pz + M[15];
    V3 nearP(M[0] * px + M[1] * py + M[2] * pz + M[3], M[4] * px + M[5] * py + M[6] * pz + M[7], M[8] * px + M[9] * py + M[10] * pz + M[11]);
    if (w != 1.0f) nearP = nearP / w;
    const V3 dx(cam.dx[0], cam.dx[1], cam.dx[2]), dy(cam.dy[0], cam.dy[1], cam.dy[2]);
    V3 lx, ly;
    if (cam.apertureRadius > 0) { // thinlens.cpp:326-357
        float tx, ty;
        squareToUniformDiskConcentric(apx, apy, tx, ty);
        const V3 apertureP(tx * cam.apertureRadius, ty * cam.apertureRadius, 0.0f);
        const float fDist = cam.focusDistance / nearP.z;
        lx = normalize((nearP + dx) * fDist - apertureP);
        ly = normalize((nearP + dy) * fDist - apertureP);
    } else { // perspective.cpp:293-294
        lx = normalize(nearP + dx);
        ly = normalize(nearP + dy);
    }
    const float *T = cam.camToWorld;
    const V3 wx(T[0] * lx.x + T[1] * lx.y + T[2] * lx.z, T[4] * lx.x + T[5] * lx.y + T[6] * lx.z, T[8] * lx.x + T[9] * lx.y + T[10] * lx.z);
    const V3 wy(T[0] * ly.x + T[1] * ly.y + T[2] * ly.z, T[4] * ly.x + T[5] * ly.y + T[6] * ly.z, T[8] * ly.x + T[9] * ly.y + T[10] * ly.z);
    rxD = d + (wx - d) * scale; // ray.h:163-168
    ryD = d + (wy - d) * scale;
}

// sampler set-up for (pixel, sample): sobol.cpp:204-216 / counter stream
B2_DEV void samplerInit(const DScene &sc, const DRender &rp, int px, int py, uint32_t s, PathSampler &smp, float &ax, float &ay) {
    smp.kind = rp.sampler;
    smp.m32 = sc.sobolNib;
    smp.nNib = rp.indexNibbles;
    smp.overflow = false; smp.cacheDim = 0xFFFFFFFFu;
    smp.dim = 0;
    if (rp.sampler == 0) {
        smp.scramble32 = (uint32_t) rp.scramble;
        uint64_t idx = s;
        if (rp.logRes > 1) idx = sobolLookUpNib(rp.lookupNib, rp.logRes, s, (uint32_t) px, (uint32_t) py, rp.scramble, rp.frameNibbles, rp.bNibbles);
        smp.index = idx;
        if (idx != (uint64_t) s) { // sobol.cpp:241-243
            ax = smp.next1D() * rp.resolution - (float) px;
            ay = smp.next1D() * rp.resolution - (float) py;
        } else {
            ax = smp.next1D();
            ay = smp.next1D();
        }
    } else {
        smp.scramble32 = (uint32_t) (rp.scramble >> 32);
        smp.index = ((((uint32_t) py * (uint32_t) sc.cam.W + (uint32_t) px) * (uint32_t) rp.spp + s) ^ (uint32_t) rp.scramble);
        ax = smp.next1D();
        ay = smp.next1D();
    }
}

// Work item -> (pixel, sample).  Items [0, tilesX*tilesY*64*nS) walk the whole 8x8 tiles of the film in rounds of roundSpp samples:
// round-major, then tile, then the sample index inside the round, then the pixel inside the tile (64 consecutive items = one sample of
// one tile: coherent camera rays and one compact film footprint).  The remaining items cover the right strip (W - 8*tilesX columns beside the tiles) and the bottom
// strip (H - 8*tilesY full rows) pixel by pixel, sample-major.  Every item is a pixel of the film, so no slot is ever wasted.
B2_DEV void workItemPixel(const DRender &rp, int W, int H, unsigned long long w, int &px, int &py, uint32_t &s) {
    const uint32_t nS = (uint32_t) (rp.sampleHi - rp.sampleLo);
    const uint32_t perTile = 64u * nS;
    const unsigned long long tiled = (unsigned long long) rp.tilesX * rp.tilesY * perTile;
    if (w < tiled) {
        // rounds of rp.roundSpp samples: every tile receives roundSpp samples, then the next round starts.  The pool then holds paths of
        // pool / (64 * roundSpp) tiles instead of pool / (64 * nS): fewer film atomics land on the same pixel at the same time (the L2
        // serialises atomics per address), while 64 consecutive items are still one sample of one tile (coherent camera rays).
        const uint32_t S = rp.roundSpp, perVisit = 64u * S;
        const unsigned long long perRound = (unsigned long long) rp.tilesX * rp.tilesY * perVisit;
        const uint32_t round = (uint32_t) (w / perRound);
        const unsigned long long rem = w % perRound;
        const uint32_t tile = (uint32_t) (rem / perVisit), r = (uint32_t) (rem % perVisit);
        s = (uint32_t) rp.sampleLo + round * S + r / 64u;
        const uint32_t p = r & 63u;
        px = (int) ((tile % rp.tilesX) * 8u + (p & 7u));
        py = (int) ((tile / rp.tilesX) * 8u + (p >> 3));
        return;
    }
    const unsigned long long e = w - tiled;
    const uint32_t Wi = rp.tilesX * 8u, Hi = rp.tilesY * 8u, rw = (uint32_t) W - Wi;
    const uint32_t nEdge = (uint32_t) W * (uint32_t) H - Wi * Hi;
    s = (uint32_t) rp.sampleLo + (uint32_t) (e / nEdge);
    uint32_t q = (uint32_t) (e % nEdge);
    if (q < rw * Hi) { px = (int) (Wi + q % rw); py = (int) (q / rw); }
    else { q -= rw * Hi; px = (int) (q % (uint32_t) W); py = (int) (Hi + q / (uint32_t) W); }
}

// ------------------------------------------------------------------------------------------------
// k_generate: drains the finished-path queue of the previous iteration with full warps: splat (ImageBlock::put),
// then refill the slot with the next (pixel, sample) work item.  FIRST: every slot is empty, no queue yet.
// The four records of a finished slot (flags, pixel, Li, sample position) are gathered through the queue index, which is fetched one
// loop iteration ahead.  B2_STAGE_GEN (A/B switch, off): also stage the records themselves one iteration ahead with cp.async --
// measured neutral on B200 (507 vs 504 ms per 5 steps): what the DRAM wait loses, the LDGSTS issue cost (MIO throttle) takes back.
// ------------------------------------------------------------------------------------------------
#define B2_GEN_BLOCK 256
#ifndef B2_GEN_MINBLOCKS
#define B2_GEN_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(B2_GEN_BLOCK, B2_GEN_MINBLOCKS) k_generate(DScene sc, DPool pool, DRender rp, DFilter filt) {
#ifdef B2_STAGE_GEN
    __shared__ __align__(16) float4 sLi[2][B2_GEN_BLOCK];
    __shared__ __align__(8) float2 sPos[2][B2_GEN_BLOCK];
    __shared__ uint32_t sFlags[2][B2_GEN_BLOCK], sPix[2][B2_GEN_BLOCK];
#endif
    const uint32_t Q = pool.capacity;
    const uint32_t it = (uint32_t) pool.counters[CTR_ITER];
    const bool FIRST = it == 0;
    stampBegin(rp, it, STAGE_GENERATE);
    const uint32_t *queue = pool.doneQueue + (size_t) ((it + 1u) & 1u) * Q; // written by k_shade of iteration it - 1
    const uint32_t n = FIRST ? Q : (uint32_t) pool.counters[((it + 1u) & 1u) ? CTR_DONE1 : CTR_DONE0];
    uint32_t nSamples = 0, nBad = 0, nNew = 0;
    uint32_t pathLen = 0;
    // work items are handed out without atomics: entry j of the drained queue takes item base + j; k_publish advances
    // CTR_NEXT by the queue length after this kernel
    const unsigned long long workBase = pool.counters[CTR_NEXT];
    const uint32_t stride = gridDim.x * blockDim.x, tid = threadIdx.x;
    auto slotOf = [&](uint32_t j) -> uint32_t { return j < n ? (FIRST ? j : __ldg(queue + j)) : 0u; };
    uint32_t j = blockIdx.x * blockDim.x + tid;
    uint32_t iCur = slotOf(j), iNext = slotOf(j + stride); // the queue entry of the next iteration is fetched one iteration ahead
#ifdef B2_STAGE_GEN
    auto stage = [&](int buf, uint32_t j, uint32_t i) {
        if (!FIRST && j < n) {
            cpAsync16(&sLi[buf][tid], &pool.st[2 * (size_t) i + 1]);
            cpAsync8(&sPos[buf][tid], &pool.pos[i]);
            cpAsync4(&sFlags[buf][tid], &pool.flags[i]);
            cpAsync4(&sPix[buf][tid], &pool.pix[i]);
        }
        cpAsyncCommit();
    };
    stage(0, j, iCur);
    int buf = 0;
#endif
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += stride, j += stride) {
        const bool inRange = j < n;
        const uint32_t i = iCur;
#ifdef B2_STAGE_GEN
        stage(buf ^ 1, j + stride, iNext);
#endif
        iCur = iNext;
        iNext = slotOf(j + 2 * stride);
#ifdef B2_STAGE_GEN
        cpAsyncWait<1>();
#endif
        if (inRange && !FIRST) {
#ifdef B2_STAGE_GEN
            const uint32_t fl = sFlags[buf][tid], pixel = sPix[buf][tid];
            const float4 li = sLi[buf][tid];
            const float2 sp = sPos[buf][tid];
#else
            const uint32_t fl = pool.flags[i], pixel = pool.pix[i];
            const float4 li = pool.st[2 * (size_t) i + 1];
            const float2 sp = pool.pos[i];
#endif
            const float alpha = (fl & PF_ALPHA) ? 1.0f : 0.0f;
            if (!filmPut(filt, rp.filmRGBA, rp.filmW, sc.cam.W, sc.cam.H, sp.x, sp.y, (int) (pixel & 0xFFFFu), (int) (pixel >> 16), V3(li.x, li.y, li.z), alpha))
                ++nBad;
            ++nSamples;
            const uint32_t len = (fl >> 8) & 0xFFFu;
            pathLen += len;
            if (rp.pixStats) atomicAdd(rp.pixStats + (size_t) (pixel >> 16) * sc.cam.W + (pixel & 0xFFFFu), ((unsigned long long) (len * len) << 32) | len);
        }
        const unsigned long long w = workBase + j;
        if (inRange) {
            const bool valid = w < rp.totalWork;
            int px = 0, py = 0;
            uint32_t s = 0;
            if (valid) workItemPixel(rp, sc.cam.W, sc.cam.H, w, px, py, s);
            if (valid) {
                PathSampler smp;
                float ax, ay;
                samplerInit(sc, rp, px, py, s, smp, ax, ay);
                const float spx = (float) px + ax, spy = (float) py + ay;
                V3 o, d;
                float mint, maxt;
                float apx = 0.5f, apy = 0.5f;
                if (sc.cam.apertureRadius > 0) smp.next2D(apx, apy); // needsApertureSample, integrator.cpp:173-174
                cameraRay(sc.cam, spx, spy, o, d, mint, maxt, apx, apy);
                pool.ray[2 * (size_t) i] = make_float4(o.x, o.y, o.z, mint);
                pool.ray[2 * (size_t) i + 1] = make_float4(d.x, d.y, d.z, maxt);
                pool.st[2 * (size_t) i] = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
                pool.st[2 * (size_t) i + 1] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                pool.smp[i] = make_uint2((uint32_t) smp.index, (uint32_t) (smp.index >> 32));
                pool.pos[i] = make_float2(spx, spy);
                pool.pix[i] = ((uint32_t) py << 16) | (uint32_t) px;
                pool.flags[i] = (PF_ALIVE | PF_FRESH) | (1u << 8) | (smp.dim << 20); // depth = 1 (integrator.h:221-227)
                ++nNew;
            } else {
                pool.flags[i] = 0; // no work left: the slot stays empty for the rest of the render
            }
        }
#ifdef B2_STAGE_GEN
        buf ^= 1;
#endif
    }
#ifdef B2_STAGE_GEN
    cpAsyncWait<0>();
#endif
    nNew = warpSum(nNew);
    nSamples = warpSum(nSamples);
    nBad = warpSum(nBad);
    pathLen = warpSum(pathLen);
    if ((threadIdx.x & 31) == 0) {
        if (nNew) atomicAdd(pool.counters + CTR_ACTIVE, (unsigned long long) nNew);
        if (nSamples) atomicAdd(pool.counters + CTR_SAMPLES, (unsigned long long) nSamples);
        if (nBad) atomicAdd(pool.counters + CTR_BAD, (unsigned long long) nBad);
        if (pathLen) atomicAdd(pool.counters + CTR_PATHLEN, (unsigned long long) pathLen);
    }
    stampEnd(rp, it, STAGE_GENERATE);
}

// counter of class queue c (the fifth one lives outside the sliding window of CTR_CLASS0..3)
B2_DEV int classCounter(int c) { return c < 4 ? CTR_CLASS0 + c : CTR_CLASSG; }

// One thread, between k_generate and k_extend: publishes progress to the host ring, clears the per-iteration queue
// counters and advances the iteration.  Keeping the iteration on the device lets ONE CUDA graph serve every iteration.
__global__ void k_publish(DPool pool, DRender rp) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    unsigned long long *c = pool.counters;
    const unsigned long long it = c[CTR_ITER];
    volatile unsigned long long *slot = rp.ring + (it % B2_RING) * 4;
    slot[1] = c[CTR_ACTIVE];
    slot[2] = c[CTR_NEXT] + (it == 0 ? (unsigned long long) pool.capacity : c[((it + 1) & 1) ? CTR_DONE1 : CTR_DONE0]);
    __threadfence_system();
    slot[0] = it + 1;
    c[CTR_TICKET_EXT] = 0; c[CTR_TICKET_OCC] = 0;
    c[CTR_SHADOW] = 0;
    c[CTR_CLASS0] = 0; c[CTR_CLASS0 + 1] = 0; c[CTR_CLASS0 + 2] = 0; c[CTR_CLASS0 + 3] = 0; c[CTR_CLASSG] = 0;
    const int dq = ((it + 1) & 1) ? CTR_DONE1 : CTR_DONE0;
    c[CTR_NEXT] += it == 0 ? (unsigned long long) pool.capacity : c[dq]; // work items k_generate just handed out
    c[dq] = 0; // drained by this iteration's k_generate, refilled by k_shade of it + 1
    c[CTR_ITER] = it + 1;
}

// ------------------------------------------------------------------------------------------------
// k_extend: closest hit for every live slot (+ optional material-class binning)
// ------------------------------------------------------------------------------------------------
template <bool SORT> __global__ void __launch_bounds__(B2_TRACE_BLOCK) k_extend(DScene sc, DPool pool, DRender rp) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t it = (uint32_t) pool.counters[CTR_ITER] - 1u; // k_publish already advanced the counter
    stampBegin(rp, it, STAGE_EXTEND);
    const bool wide = sc.nodes8 != nullptr && !sc.rootCount && !sc.nItems;
    const TraceMem tm = setupTraceMem(sc, smem, wide);
    const uint32_t Q = pool.capacity;
    uint32_t nRays = 0;
    if (!sc.rootCount && !sc.nItems) {
        // BVH scenes: persistent loop with per-lane ray replacement (b2_trace.cuh: traverseQueue)
        uint32_t nv = 0, pt = 0;
        auto fetch = [&](uint32_t i, V3 &o, V3 &d, float &mint, float &maxt) -> int {
            const uint32_t fl = pool.flags[i];
            if (!(fl & PF_ALIVE)) return 0;
            const float4 ro = pool.ray[2 * (size_t) i];
            float4 rd = pool.ray[2 * (size_t) i + 1];
            if (!(fl & PF_FRESH)) rd.w = B2_INF;
            o = V3(ro.x, ro.y, ro.z); d = V3(rd.x, rd.y, rd.z);
            const V3 dRcp(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
            ++nRays;
            // non-camera rays: mint = Epsilon (ray.h:66-68); their ray[2i].w carries dot(wo, refN) for the environment MIS term
            return clipRay<false>(sc, o, d, dRcp, (fl & PF_FRESH) ? ro.w : B2_EPSILON, rd.w, mint, maxt) ? 2 : 1;
        };
        auto commit = [&](uint32_t i, bool found, const HitRec &h) {
            pool.hit[i] = found ? make_float4(h.t, h.u, h.v, __uint_as_float(h.prim)) : make_float4(B2_INF, 0.0f, 0.0f, __uint_as_float(0xFFFFFFFFu));
            if (SORT) {
                int cls = (int) sc.missClass;
                if (found) cls = min((int) sc.materials[__float_as_int(__ldg(&sc.verts[3 * (size_t) h.prim].w))].type, B2_NCLASS - 1);
#pragma unroll
                for (int c = 0; c < B2_NCLASS; ++c) {
                    const uint32_t at = warpAppend(cls == c, pool.counters + classCounter(c));
                    if (cls == c) pool.matQueue[(size_t) c * Q + at] = i;
                }
            }
        };
        if (wide) traverseQueue8<false, false>(sc, tm, Q, pool.counters + CTR_TICKET_EXT, fetch, commit, nv, pt);
        else traverseQueue<false, false>(sc, tm, Q, pool.counters + CTR_TICKET_EXT, fetch, commit, nv, pt);
    } else
    for (uint32_t base = blockIdx.x * blockDim.x; base < Q; base += gridDim.x * blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t fl = i < Q ? pool.flags[i] : 0u;
        const bool live = (fl & PF_ALIVE) != 0;
        int cls = -1;
        if (live) {
            const float4 ro = pool.ray[2 * (size_t) i];
            float4 rd = pool.ray[2 * (size_t) i + 1];
            if (!(fl & PF_FRESH)) rd.w = B2_INF; // w carries the pending BSDF pdf for non-camera rays; their maxt is +inf (ray.h:66-68)
            const V3 o(ro.x, ro.y, ro.z), d(rd.x, rd.y, rd.z);
            const V3 dRcp(1.0f / d.x, 1.0f / d.y, 1.0f / d.z); // ray.h:83-84
            HitRec h;
            h.t = B2_INF; h.u = 0; h.v = 0; h.prim = 0xFFFFFFFFu;
            float mint, maxt;
            uint32_t nv = 0, pt = 0;
            uint32_t item = 0xFFFFFFFFu;
            if (clipRay<false>(sc, o, d, dRcp, (fl & PF_FRESH) ? ro.w : B2_EPSILON, rd.w, mint, maxt)) {
                const bool found = sc.nItems ? traverseTop<false, false>(sc, tm, o, d, mint, maxt, h, item, nv, pt) : traverse<false, false>(sc, tm, o, d, mint, maxt, h, nv, pt);
                if (!found) { h.t = B2_INF; h.prim = 0xFFFFFFFFu; }
            }
            if (sc.nItems) pool.inst[i] = item;
            pool.hit[i] = make_float4(h.t, h.u, h.v, __uint_as_float(h.prim));
            ++nRays;
            if (SORT) {
                cls = (int) sc.missClass;
                if (h.prim != 0xFFFFFFFFu) {
                    const int mat = __float_as_int(__ldg(&sc.verts[3 * (size_t) h.prim].w));
                    cls = min((int) sc.materials[mat].type, B2_NCLASS - 1);
                }
            }
        }
        if (SORT) {
#pragma unroll
            for (int c = 0; c < B2_NCLASS; ++c) {
                const uint32_t at = warpAppend(cls == c, pool.counters + classCounter(c));
                if (cls == c) pool.matQueue[(size_t) c * Q + at] = i;
            }
        }
    }
    nRays = warpSum(nRays);
    if ((threadIdx.x & 31) == 0 && nRays) atomicAdd(pool.counters + CTR_RAYS, (unsigned long long) nRays);
    stampEnd(rp, it, STAGE_EXTEND);
}

// ------------------------------------------------------------------------------------------------
// Flat-leaf variants (DScene::rootCount > 0: the whole scene is one shared-memory resident leaf, e.g. the Cornell box): own kernels so
// that the register allocation is not the maximum over the BVH / instanced code paths (72 -> ~56 registers: 4 instead of 3 CTAs per
// SM), no traversal stack in shared memory, and the records of the NEXT item of the grid-stride loop are loaded into registers before
// the current ray is tested (ncu, round 2: a quarter of these kernels' stall samples sat on the first use of the freshly loaded ray).
// ------------------------------------------------------------------------------------------------
B2_DEV TraceMem setupFlatLeaf(const DScene &sc, unsigned char *smem) {
    float4 *sTris = (float4 *) smem;
    const size_t off = ((size_t) sc.stageTriBytes + 15) & ~(size_t) 15;
    uint64_t *bar = (uint64_t *) (smem + off);
    DScene tmp = sc;
    tmp.stageNodes = 0; // the flat leaf has no node array
    stageScene(tmp, sTris, sTris, bar);
    TraceMem tm;
    tm.gNodes = nullptr; tm.sNodes = nullptr; tm.stageNodes = 0;
    tm.sTris = sTris; tm.stageTris = sc.stageTris; tm.stack = nullptr; tm.stride = 0;
#ifdef B2_FAST_TRI
    tm.gTris = sc.triPlane;
#else
    tm.gTris = sc.triAccel;
#endif
    return tm;
}

template <bool SORT> __global__ void __launch_bounds__(B2_TRACE_BLOCK, 4) k_extend_flat(DScene sc, DPool pool, DRender rp) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t it = (uint32_t) pool.counters[CTR_ITER] - 1u;
    stampBegin(rp, it, STAGE_EXTEND);
    const TraceMem tm = setupFlatLeaf(sc, smem);
    const uint32_t Q = pool.capacity;
    uint32_t nRays = 0;
    const uint32_t stride = gridDim.x * blockDim.x;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t flN = 0;
    float4 roN = make_float4(0, 0, 0, 0), rdN = roN;
    if (i < Q) { flN = pool.flags[i]; roN = pool.ray[2 * (size_t) i]; rdN = pool.ray[2 * (size_t) i + 1]; }
    for (uint32_t base = blockIdx.x * blockDim.x; base < Q; base += stride, i += stride) {
        const uint32_t fl = flN;
        const float4 ro = roN;
        float4 rd = rdN;
        const uint32_t iN = i + stride;
        if (iN < Q) { flN = pool.flags[iN]; roN = pool.ray[2 * (size_t) iN]; rdN = pool.ray[2 * (size_t) iN + 1]; } else flN = 0;
        const bool live = i < Q && (fl & PF_ALIVE) != 0;
        int cls = -1;
        if (live) {
            if (!(fl & PF_FRESH)) rd.w = B2_INF; // w carries the pending BSDF pdf for non-camera rays; their maxt is +inf (ray.h:66-68)
            const V3 o(ro.x, ro.y, ro.z), d(rd.x, rd.y, rd.z);
            const V3 dRcp(1.0f / d.x, 1.0f / d.y, 1.0f / d.z); // ray.h:83-84
            HitRec h;
            h.t = B2_INF; h.u = 0; h.v = 0; h.prim = 0xFFFFFFFFu;
            float mint, maxt;
            uint32_t pt = 0;
            if (clipRay<false>(sc, o, d, dRcp, (fl & PF_FRESH) ? ro.w : B2_EPSILON, rd.w, mint, maxt))
                if (!traverseFlat<false, false>(sc, tm, o, d, mint, maxt, h, pt)) { h.t = B2_INF; h.prim = 0xFFFFFFFFu; }
            pool.hit[i] = make_float4(h.t, h.u, h.v, __uint_as_float(h.prim));
            ++nRays;
            if (SORT) {
                cls = (int) sc.missClass;
                if (h.prim != 0xFFFFFFFFu) cls = min((int) sc.materials[__float_as_int(__ldg(&sc.verts[3 * (size_t) h.prim].w))].type, B2_NCLASS - 1);
            }
        }
        if (SORT) {
#pragma unroll
            for (int c = 0; c < B2_NCLASS; ++c) {
                const uint32_t at = warpAppend(cls == c, pool.counters + classCounter(c));
                if (cls == c) pool.matQueue[(size_t) c * Q + at] = i;
            }
        }
    }
    nRays = warpSum(nRays);
    if ((threadIdx.x & 31) == 0 && nRays) atomicAdd(pool.counters + CTR_RAYS, (unsigned long long) nRays);
    stampEnd(rp, it, STAGE_EXTEND);
}

__global__ void __launch_bounds__(B2_TRACE_BLOCK, 4) k_occluded_flat(DScene sc, DPool pool, DRender rp) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t it = (uint32_t) pool.counters[CTR_ITER] - 1u;
    stampBegin(rp, it, STAGE_OCCLUDED);
    const TraceMem tm = setupFlatLeaf(sc, smem);
    const uint32_t n = (uint32_t) pool.counters[CTR_SHADOW];
    uint32_t nClear = 0;
    const uint32_t stride = gridDim.x * blockDim.x;
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    float4 roN = make_float4(0, 0, 0, 0), sdN = roN, scN = roN;
    if (j < n) { roN = pool.shO[j]; sdN = pool.shD[j]; scN = pool.shC[j]; }
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += stride, j += stride) {
        const float4 ro = roN, sd = sdN, scn = scN;
        const uint32_t jN = j + stride;
        if (jN < n) { roN = pool.shO[jN]; sdN = pool.shD[jN]; scN = pool.shC[jN]; }
        if (j < n) {
            const V3 o(ro.x, ro.y, ro.z), d(sd.x, sd.y, sd.z);
            const V3 dRcp(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
            float mint, maxt;
            bool occluded = false;
            HitRec h;
            uint32_t pt = 0;
            if (clipRay<true>(sc, o, d, dRcp, B2_EPSILON, sd.w, mint, maxt)) occluded = traverseFlat<true, false>(sc, tm, o, d, mint, maxt, h, pt);
            if (!occluded) { // Li += value * bsdfVal * weight (path.cpp:197): reduction at L2, no read of the slot
                redAddV4(&pool.st[2 * (size_t) __float_as_uint(scn.w) + 1], scn.x, scn.y, scn.z, 0.0f);
                ++nClear;
            }
        }
    }
    nClear = warpSum(nClear);
    if ((threadIdx.x & 31) == 0 && nClear) atomicAdd(pool.counters + CTR_UNOCCLUDED, (unsigned long long) nClear);
    stampEnd(rp, it, STAGE_OCCLUDED);
}

// ------------------------------------------------------------------------------------------------
// intersection record: skdtree.h:343-428 fillIntersectionRecord<true>
// ------------------------------------------------------------------------------------------------
struct Isect {
    V3 p;
    V3 geoN;
    Frame sh;
    V3 wi;
    int material, emitter;
};
B2_DEV void fillIntersection(const DScene &sc, const V3 &rayD, uint32_t prim, float u, float v, Isect &its) {
    const float4 a0 = __ldg(&sc.verts[3 * (size_t) prim]), a1 = __ldg(&sc.verts[3 * (size_t) prim + 1]), a2 = __ldg(&sc.verts[3 * (size_t) prim + 2]);
    const V3 p0(a0.x, a0.y, a0.z), p1(a1.x, a1.y, a1.z), p2(a2.x, a2.y, a2.z);
    its.material = __float_as_int(a0.w);
    its.emitter = __float_as_int(a1.w);
    const uint32_t tflags = __float_as_uint(a2.w);
    const V3 b(1 - u - v, u, v);
    its.p = p0 * b.x + p1 * b.y + p2 * b.z;
    const V3 side1 = p1 - p0, side2 = p2 - p0;
    V3 faceNormal = cross(side1, side2);
    const float len = length(faceNormal);
    if (!isZero(faceNormal)) faceNormal = faceNormal / len;
    V3 dpdu = side1;
    V3 shN;
    if (tflags & 3u) {
        const float4 n0 = __ldg(&sc.norms[3 * (size_t) prim]), n1 = __ldg(&sc.norms[3 * (size_t) prim + 1]), n2 = __ldg(&sc.norms[3 * (size_t) prim + 2]);
        if (tflags & 2u) dpdu = V3(n0.w, n1.w, n2.w);
        if (tflags & 1u) {
            shN = normalize(V3(n0.x, n0.y, n0.z) * b.x + V3(n1.x, n1.y, n1.z) * b.y + V3(n2.x, n2.y, n2.z) * b.z);
            if (dot(faceNormal, shN) < 0) faceNormal = -faceNormal;
        } else shN = faceNormal;
    } else shN = faceNormal;
    its.geoN = faceNormal;
    computeShadingFrame(shN, dpdu, its.sh);
    its.wi = its.sh.toLocal(-rayD);
}

// Instance::fillIntersectionRecord (instance.cpp:149-162): the nested record is filled in object space with p = ray'(t)
// (skdtree.h fillIntersectionRecord<false>), then normals (inverse transpose), dpdu and p go to world space and the shading
// frame is rebuilt from the transformed quantities (skdtree.h:424-427)
B2_DEV void fillIntersectionInst(const DScene &sc, const V3 &rayO, const V3 &rayD, float t, uint32_t prim, float u, float v, const DInstance &in, Isect &its) {
    const V3 oo = xfPoint(in.Minv, rayO), od = xfVector(in.Minv, rayD);
    const float4 a0 = __ldg(&sc.verts[3 * (size_t) prim]), a1 = __ldg(&sc.verts[3 * (size_t) prim + 1]), a2 = __ldg(&sc.verts[3 * (size_t) prim + 2]);
    const V3 p0(a0.x, a0.y, a0.z), p1(a1.x, a1.y, a1.z), p2(a2.x, a2.y, a2.z);
    its.material = __float_as_int(a0.w);
    its.emitter = __float_as_int(a1.w);
    const uint32_t tflags = __float_as_uint(a2.w);
    const V3 b(1 - u - v, u, v);
    const V3 side1 = p1 - p0, side2 = p2 - p0;
    V3 faceNormal = cross(side1, side2);
    const float len = length(faceNormal);
    if (!isZero(faceNormal)) faceNormal = faceNormal / len;
    V3 dpdu = side1;
    V3 shN;
    if (tflags & 3u) {
        const float4 n0 = __ldg(&sc.norms[3 * (size_t) prim]), n1 = __ldg(&sc.norms[3 * (size_t) prim + 1]), n2 = __ldg(&sc.norms[3 * (size_t) prim + 2]);
        if (tflags & 2u) dpdu = V3(n0.w, n1.w, n2.w);
        if (tflags & 1u) {
            shN = normalize(V3(n0.x, n0.y, n0.z) * b.x + V3(n1.x, n1.y, n1.z) * b.y + V3(n2.x, n2.y, n2.z) * b.z);
            if (dot(faceNormal, shN) < 0) faceNormal = -faceNormal;
        } else shN = faceNormal;
    } else shN = faceNormal;
    its.p = xfPoint(in.M, oo + od * t);
    its.geoN = normalize(xfNormal(in.Minv, faceNormal));
    computeShadingFrame(normalize(xfNormal(in.Minv, shN)), xfVector(in.M, dpdu), its.sh);
    its.wi = its.sh.toLocal(-rayD);
}

// Texture look-up of one intersection (`m_reflectance->eval(its)` of the diffuse leaf, diffuse.cpp:115,148): uv from the per-prim texture
// coordinates (skdtree.h:398-405), and for a camera ray the uv partials of Intersection::computePartials from the sensor's ray
// differentials (perspective.cpp:271-298 / thinlens.cpp:324-361, scaled by 1/sqrt(spp): integrator.cpp:144-145,181), which this
// recomputes from the film position instead of carrying them in the pool.  `in`: the instance that was hit or null.
B2_DEV TexCoord surfaceTexCoord(const DScene &sc, float diffScale, uint32_t prim, float bu, float bv, const DInstance *in, const V3 &p, const V3 &geoN,
                                bool fresh, float2 pos, PathSampler smp) {
    const size_t p3 = 3 * (size_t) prim;
    const float4 a0 = __ldg(&sc.verts[p3]), a1 = __ldg(&sc.verts[p3 + 1]), a2 = __ldg(&sc.verts[p3 + 2]);
    const uint32_t tflags = __float_as_uint(a2.w);
    TexCoord tc;
    tc.hasUVPartials = false;
    tc.dudx = tc.dudy = tc.dvdx = tc.dvdy = 0.0f;
    V3 dpdu, dpdv;
    if (tflags & 2u) { // the mesh has texture coordinates: interpolated uv, tangents of TriMesh::computeUVTangents
        const float4 t0 = __ldg(&sc.texc[p3]), t1 = __ldg(&sc.texc[p3 + 1]), t2 = __ldg(&sc.texc[p3 + 2]);
        const float b0 = 1 - bu - bv;
        tc.u = t0.x * b0 + t1.x * bu + t2.x * bv;
        tc.v = t0.y * b0 + t1.y * bu + t2.y * bv;
        dpdu = V3(__ldg(&sc.norms[p3].w), __ldg(&sc.norms[p3 + 1].w), __ldg(&sc.norms[p3 + 2].w));
        dpdv = V3(t0.z, t1.z, t2.z);
    } else { // skdtree.h:377-379,403-404
        tc.u = bu; tc.v = bv;
        const V3 p0(a0.x, a0.y, a0.z);
        dpdu = V3(a1.x, a1.y, a1.z) - p0;
        dpdv = V3(a2.x, a2.y, a2.z) - p0;
    }
    if (fresh) {
        if (in) { dpdu = xfVector(in->M, dpdu); dpdv = xfVector(in->M, dpdv); } // instance.cpp:158-159
        float apx = 0.5f, apy = 0.5f;
        if (sc.cam.apertureRadius > 0) { smp.dim = 2; smp.next2D(apx, apy); } // the aperture sample of this path (dimensions 2, 3)
        V3 o, d, rxD, ryD;
        cameraRayDifferential(sc.cam, pos.x, pos.y, apx, apy, diffScale, o, d, rxD, ryD);
        computeUVPartials(p, geoN, dpdu, dpdv, o, rxD, ryD, tc);
    }
    return tc;
}
static __device__ __noinline__ Spectrum shadeTexture(const DScene &sc, const DRender &rp, int tex, uint32_t prim, float bu, float bv, const DInstance *in,
                                                   const Isect &its, bool fresh, float2 pos, PathSampler smp) {
    const TexCoord tc = surfaceTexCoord(sc, rp.diffScale, prim, bu, bv, in, its.p, its.geoN, fresh, pos, smp);
    return texEval(sc.textures[tex], sc.ewaLut, tc);
}

// Scene::evalEnvironment for a sensor ray that left the scene (path.cpp:136-143, volpath.cpp:190-202): the filtered look-up of
// envmap.cpp:392-406 with the sensor's ray differentials, recomputed from the film position like the texture partials above
static __device__ __noinline__ Spectrum envEvalCamera(const DScene &sc, float diffScale, float2 pos, PathSampler smp) {
    float apx = 0.5f, apy = 0.5f;
    if (sc.cam.apertureRadius > 0) { smp.dim = 2; smp.next2D(apx, apy); }
    V3 o, d, rxD, ryD;
    cameraRayDifferential(sc.cam, pos.x, pos.y, apx, apy, diffScale, o, d, rxD, ryD);
    return envEval(*sc.envmap, sc.ewaLut, d, true, rxD, ryD);
}

B2_DEV float miWeight(float pdfA, float pdfB) { // path.cpp:296-300
    pdfA *= pdfA;
    pdfB *= pdfB;
    return pdfA / (pdfA + pdfB);
}

// DiscreteDistribution::sample (pmf.h:128-141): lower_bound on the cdf, index clamp, zero-probability skip
B2_DEV uint32_t cdfSample(const float *__restrict__ cdf, uint32_t n /* entries, cdf has n+1 */, float v) {
    uint32_t lo = 0, cnt = n + 1; // lower_bound over cdf[0..n]
    while (cnt > 0) {
        uint32_t step = cnt >> 1, it = lo + step;
        if (__ldg(cdf + it) < v) { lo = it + 1; cnt -= step + 1; }
        else cnt = step;
    }
    int idx = (int) lo - 1;
    if (idx < 0) idx = 0;
    if ((uint32_t) idx > n - 1) idx = (int) n - 1;
    while ((__ldg(cdf + idx + 1) - __ldg(cdf + idx)) == 0 && (uint32_t) idx < n) ++idx;
    return (uint32_t) idx;
}

struct DirectSample {
    V3 d, p, n;
    float dist, pdf;
    Spectrum value;
    int emitter;
};
// Scene::sampleEmitterDirect (scene.cpp:828-852) up to, not including, the visibility ray:
// emitter pick (pmf.h sampleReuse), AreaLight::sampleDirect (area.cpp:158-173), Shape::sampleDirect
// (shape.cpp:102-115), TriMesh::samplePosition (trimesh.cpp:412-424), Triangle::sample (triangle.cpp:24-62)
// withMap: compile-time switch of the call site -- the untextured k_shade instances leave the environment-map code out (scenes with a
// map are shaded by the textured instance, see launch_shade), so their registers and stack stay what they were without it
B2_DEV bool sampleEmitterDirect(const DScene &sc, const V3 &ref, const V3 &refN, float sx, float sy, DirectSample &ds, float *emPdfOut = nullptr,
                                const bool withMap = true) {
    const uint32_t ei = cdfSample(sc.emitterCdf, sc.nEmitters, sx);
    const float c0 = __ldg(sc.emitterCdf + ei), c1 = __ldg(sc.emitterCdf + ei + 1);
    const float emPdf = c1 - c0;
    sx = (sx - c0) / (c1 - c0);
    const DEmitter &em = sc.emitters[ei];
    if (em.nTri == 0) { // environment emitter: `constant` (constant.cpp:171-208) or `envmap` (envmap.cpp:516-543)
        V3 d;
        float pdf;
        Spectrum rad(em.radiance[0], em.radiance[1], em.radiance[2]);
        const bool isMap = withMap && sc.envmap != nullptr;
        if (isMap) {
            V3 dl;
            envSampleDirection(*sc.envmap, sx, sy, dl, rad, pdf);
            d = envToWorld(*sc.envmap, dl);
        } else if (!isZero(refN)) {
            d = squareToCosineHemisphere(sx, sy);
            pdf = squareToCosineHemispherePdf(d);
            Frame f;
            f.n = refN;
            coordinateSystem(f.n, f.s, f.t);
            d = f.toWorld(d);
        } else {
            const float z = 1.0f - 2.0f * sy, r = safe_sqrt(1.0f - z * z); // warp.cpp:25-31
            float sinPhi, cosPhi;
            sincosf(2.0f * B2_PI * sx, &sinPhi, &cosPhi);
            d = V3(r * cosPhi, r * sinPhi, z);
            pdf = 0.07957747154594766788f;
        }
        ds.pdf = 0.0f; ds.value = Spectrum(0.0f); ds.emitter = (int) ei;
        if (isMap && (isZero(rad) || pdf == 0)) return false;
        // bsphere.h:88-95 + util.cpp:447-485
        const V3 c(sc.bsCenter[0], sc.bsCenter[1], sc.bsCenter[2]);
        const V3 o = ref - c;
        const float A = lengthSquared(d), B = 2 * dot(o, d), C = lengthSquared(o) - sc.bsRadius * sc.bsRadius;
        const float discrim = B * B - 4.0f * A * C;
        if (A == 0 || discrim < 0) return false;
        const float sq = sqrtf(discrim), temp = B < 0 ? -0.5f * (B - sq) : -0.5f * (B + sq);
        float x0 = temp / A, x1 = C / temp;
        if (x0 > x1) { const float t = x0; x0 = x1; x1 = t; }
        if (!(x0 < 0 && x1 > 0)) return false;
        ds.p = ref + d * x1;
        ds.n = normalize(c - ds.p);
        ds.d = d; ds.dist = x1;
        if (!isMap && !isZero(refN) && dot(d, refN) <= 0) return false; // constant.cpp: roundoff moved the sample to the back side: value 0
        ds.value = rad / pdf;
        if (emPdfOut) { ds.pdf = pdf; *emPdfOut = emPdf; return true; }
        ds.pdf = pdf * emPdf;
        ds.value = ds.value / emPdf;
        return true;
    }
    const float *tcdf = sc.triCdf + em.cdfOffset;
    const uint32_t ti = cdfSample(tcdf, em.nTri, sy);
    const float t0 = __ldg(tcdf + ti), t1 = __ldg(tcdf + ti + 1);
    sy = (sy - t0) / (t1 - t0);
    const size_t prim = (size_t) em.primOffset + ti;
    const float4 a0 = __ldg(&sc.verts[3 * prim]), a1 = __ldg(&sc.verts[3 * prim + 1]), a2 = __ldg(&sc.verts[3 * prim + 2]);
    const V3 p0(a0.x, a0.y, a0.z), p1(a1.x, a1.y, a1.z), p2(a2.x, a2.y, a2.z);
    float bx, by;
    squareToUniformTriangle(sx, sy, bx, by);
    const V3 sideA = p1 - p0, sideB = p2 - p0;
    ds.p = p0 + (sideA * bx) + (sideB * by);
    if (__float_as_uint(a2.w) & 1u) {
        const float4 n0 = __ldg(&sc.norms[3 * prim]), n1 = __ldg(&sc.norms[3 * prim + 1]), n2 = __ldg(&sc.norms[3 * prim + 2]);
        ds.n = normalize(V3(n0.x, n0.y, n0.z) * (1.0f - bx - by) + V3(n1.x, n1.y, n1.z) * bx + V3(n2.x, n2.y, n2.z) * by);
    } else ds.n = normalize(cross(sideA, sideB));
    float pdf = em.invSurfaceArea;
    ds.d = ds.p - ref;
    const float distSquared = lengthSquared(ds.d);
    ds.dist = sqrtf(distSquared);
    ds.d = ds.d / ds.dist;
    const float dp = absDot(ds.d, ds.n);
    pdf *= dp != 0 ? (distSquared / dp) : 0.0f;
    ds.emitter = (int) ei;
    if (dot(ds.d, refN) >= 0 && dot(ds.d, ds.n) < 0 && pdf != 0) {
        ds.value = V3(em.radiance[0], em.radiance[1], em.radiance[2]) / pdf;
        if (emPdfOut) { // sampleAttenuatedEmitterDirect (scene.cpp:854-898): the caller multiplies by transmittance / emPdf
            ds.pdf = pdf;
            *emPdfOut = emPdf;
            return true;
        }
        ds.pdf = pdf * emPdf;
        ds.value = ds.value / emPdf;
        return true;
    }
    ds.pdf = 0.0f;
    ds.value = Spectrum(0.0f);
    return false;
}

// ------------------------------------------------------------------------------------------------
// k_shade
// ------------------------------------------------------------------------------------------------
// resident CTAs per SM the register allocator must leave room for.  Measured on B200 (round 2, Cornell 1024^2 @ 1024 spp, ms of k_shade
// per step): 4 -> 303, 5 -> 277, 6 -> 289, 8 -> 328.  5 x 128 threads leaves 102 registers: no spills, 20 warps per SM.
#ifndef B2_SHADE_MINBLOCKS
#define B2_SHADE_MINBLOCKS 5
#endif
// B2_STAGE_SHADE (A/B switch, off): stage the five records every item needs (flags, hit, ray direction + pdf, throughput + eta, sampler
// state: 60 bytes) one loop iteration ahead with cp.async.  Measured on B200 (Cornell 1024^2 @ 1024 spp): 8 % SLOWER -- the kernel is
// bound by the dependent chain of the shading body at 6 warps per scheduler, not by these loads; an LDGSTS costs 8 issue cycles.
template <int CLS, bool TEX = false> __global__ void __launch_bounds__(B2_SHADE_BLOCK, B2_SHADE_MINBLOCKS) k_shade(DScene sc, DPool pool, DRender rp, const uint32_t *queue,
                                                                             const unsigned long long *queueCount) {
#ifdef B2_STAGE_SHADE
    __shared__ __align__(16) float4 sHit[2][B2_SHADE_BLOCK], sRd[2][B2_SHADE_BLOCK], sThr[2][B2_SHADE_BLOCK];
    __shared__ __align__(8) uint2 sSmp[2][B2_SHADE_BLOCK];
    __shared__ uint32_t sState[2][B2_SHADE_BLOCK];
#endif
    const uint32_t Q = pool.capacity;
    const uint32_t n = queue ? (uint32_t) *queueCount : Q;
    const uint32_t it = (uint32_t) pool.counters[CTR_ITER] - 1u;
    stampBegin(rp, it, STAGE_SHADE);
    uint32_t nDimOvf = 0, nShadowRef = 0, nDone = 0;
    const uint32_t stride = gridDim.x * blockDim.x, tid = threadIdx.x;
    // queue appends in flight (see the end of the loop body): ballots of the lanes that append, queue bases (lane 0), parked records
    __shared__ __align__(16) float4 sShO[B2_SHADE_BLOCK], sShD[B2_SHADE_BLOCK], sShC[B2_SHADE_BLOCK];
    __shared__ uint32_t sDoneSlot[B2_SHADE_BLOCK];
    uint32_t pendA = 0, pendB = 0, baseA = 0, baseB = 0;
    auto flushPending = [&]() {
        if ((pendA | pendB) == 0) return; // warp-uniform
        const uint32_t lane = tid & 31u, below = (1u << lane) - 1u;
        const uint32_t bA = __shfl_sync(0xffffffffu, baseA, 0), bB = __shfl_sync(0xffffffffu, baseB, 0);
        if ((pendA >> lane) & 1u) pool.doneQueue[(size_t) (it & 1u) * Q + bA + __popc(pendA & below)] = sDoneSlot[tid];
        if ((pendB >> lane) & 1u) {
            const uint32_t at = bB + __popc(pendB & below);
            pool.shO[at] = sShO[tid]; pool.shD[at] = sShD[tid]; pool.shC[at] = sShC[tid];
        }
        pendA = pendB = 0;
    };
#ifdef B2_STAGE_SHADE
    auto slotOf = [&](uint32_t j) -> uint32_t { return j < n ? (queue ? __ldg(queue + j) : j) : 0u; };
    auto stage = [&](int buf, uint32_t j, uint32_t i) {
        if (j < n) {
            cpAsync4(&sState[buf][tid], &pool.flags[i]);
            cpAsync16(&sHit[buf][tid], &pool.hit[i]);
            cpAsync16(&sRd[buf][tid], &pool.ray[2 * (size_t) i + 1]);
            cpAsync16(&sThr[buf][tid], &pool.st[2 * (size_t) i]);
            cpAsync8(&sSmp[buf][tid], &pool.smp[i]);
        }
        cpAsyncCommit();
    };
    uint32_t iCur = slotOf(blockIdx.x * blockDim.x + tid), iNext = slotOf(blockIdx.x * blockDim.x + tid + stride);
    stage(0, blockIdx.x * blockDim.x + tid, iCur);
    int buf = 0;
#endif
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += stride) {
        const uint32_t j = base + threadIdx.x;
        bool live = j < n;
        uint32_t state = 0;
        float4 hit = make_float4(0, 0, 0, 0), rd4 = hit, thr4 = hit;
        uint2 sm2 = make_uint2(0, 0);
#ifdef B2_STAGE_SHADE
        const uint32_t i = iCur;
        stage(buf ^ 1, j + stride, iNext);
        iCur = iNext;
        iNext = slotOf(j + 2 * stride);
        cpAsyncWait<1>();
        if (live) {
            state = sState[buf][tid];
            hit = sHit[buf][tid];
            rd4 = sRd[buf][tid];
            thr4 = sThr[buf][tid];
            sm2 = sSmp[buf][tid];
        }
        buf ^= 1;
#else
        uint32_t i = 0;
        if (live) i = queue ? queue[j] : j;
        // all loads of the slot are issued together (memory-level parallelism); Li is never read here (it is updated by reductions)
        if (live) {
            state = pool.flags[i];
            hit = pool.hit[i];
            rd4 = pool.ray[2 * (size_t) i + 1];
            thr4 = pool.st[2 * (size_t) i];
            sm2 = pool.smp[i];
        }
#endif
        uint32_t flags = state & 0xFFu;
        if (!queue) live = live && (flags & PF_ALIVE);
        // shadow-ray output of this lane
        bool emitShadow = false;
        V3 shD(0.0f), shO(0.0f);
        float shMaxt = 0;
        Spectrum shC(0.0f);
        if (live) {
            int depth = (int) ((state >> 8) & 0xFFFu);
            const int vertex = (state & PF_FRESH) ? 1 : depth + 1; // (diagnostics) the path vertex this invocation shades
            uint32_t endKind = 0;
            bool transmitted = false;
            const V3 rayD(rd4.x, rd4.y, rd4.z);
            Spectrum T(thr4.x, thr4.y, thr4.z), LiAdd(0.0f); // radiance added by this invocation (emitter / environment hit)
            bool liTouched = false;
            float eta = thr4.w;
            const float bsdfPdfPrev = rd4.w;
            float bsdfPdfOut = 0.0f;
            PathSampler smp;
            smp.kind = rp.sampler;
            smp.m32 = sc.sobolNib;
            smp.nNib = rp.indexNibbles;
            smp.overflow = false; smp.cacheDim = 0xFFFFFFFFu;
            smp.index = ((uint64_t) sm2.y << 32) | sm2.x;
            smp.dim = state >> 20;
            smp.scramble32 = rp.sampler == 0 ? (uint32_t) rp.scramble : (uint32_t) (rp.scramble >> 32);
            const uint32_t prim = __float_as_uint(hit.w);
            const bool valid = prim != 0xFFFFFFFFu;
            bool done = false;
            Isect its;
            if (valid) {
                const uint32_t item = sc.nItems ? pool.inst[i] : 0xFFFFFFFFu;
                if (item != 0xFFFFFFFFu && !sc.items[item].identity) {
                    const float4 ro4 = pool.ray[2 * (size_t) i];
                    fillIntersectionInst(sc, V3(ro4.x, ro4.y, ro4.z), rayD, hit.x, prim, hit.y, hit.z, sc.items[item], its);
                } else fillIntersection(sc, rayD, prim, hit.y, hit.z, its);
            }
            const bool fresh = (flags & PF_FRESH) != 0;
            if (fresh) {
                if (valid) flags |= PF_ALPHA; // records.inl:117-144 (EOpacity)
            } else {
                // ---- tail of the previous loop iteration: path.cpp:226-286 ----
                if (!valid) { // :239-252: the BSDF sample left the scene
                    if (sc.envEmitter >= 0 && !(rp.hideEmitters && !(flags & PF_SCATTERED))) {
                        const DEmitter &em = sc.emitters[sc.envEmitter];
                        float lumPdf = 0.0f;
                        if (!(flags & PF_DELTA)) { // constant.cpp:210-224 / envmap.cpp:545-556 (ESolidAngle) x emitter pick probability
                            float pdfSA;
                            if (TEX && sc.envmap) pdfSA = envPdfDirection(*sc.envmap, envToLocal(*sc.envmap, rayD));
                            else {
                                const float c = pool.ray[2 * (size_t) i].w; // dot(wo, refN) of the vertex that sampled this ray; 2 = refN is zero
                                pdfSA = c == 2.0f ? 0.07957747154594766788f : B2_INV_PI * fmaxf(0.0f, c);
                            }
                            lumPdf = pdfSA * (em.samplingWeight * sc.emitterNormalization);
                        }
                        const Spectrum le = (TEX && sc.envmap) ? envEval(*sc.envmap, sc.ewaLut, rayD, false, V3(0.0f), V3(0.0f)) : V3(em.radiance[0], em.radiance[1], em.radiance[2]);
                        LiAdd = T * le * miWeight(bsdfPdfPrev, lumPdf);
                        liTouched = true;
                    }
                    done = true;
                } else {
                    if (its.emitter >= 0) {
                        const DEmitter &em = sc.emitters[its.emitter];
                        // its.Le(-ray.d): area.cpp:104-109
                        Spectrum value = dot(its.sh.n, -rayD) <= 0 ? Spectrum(0.0f) : V3(em.radiance[0], em.radiance[1], em.radiance[2]);
                        // pdfEmitterDirect: scene.cpp:949-952, area.cpp:175-183, shape.cpp:117-126
                        float lumPdf = 0.0f;
                        if (!(flags & PF_DELTA)) {
                            float pdfDirect = 0.0f;
                            if ((flags & PF_REFN_OK) && dot(rayD, its.sh.n) < 0)
                                pdfDirect = em.invSurfaceArea * (hit.x * hit.x) / absDot(rayD, its.sh.n);
                            lumPdf = pdfDirect * (em.samplingWeight * sc.emitterNormalization);
                        }
                        LiAdd = T * value * miWeight(bsdfPdfPrev, lumPdf);
                        liTouched = true;
                    }
                    // rRec.type = ERadianceNoEmission; Russian roulette :276-286
                    if (depth++ >= rp.rrDepth) {
                        float q = fminf(maxComp(T) * eta * eta, 0.95f);
                        if (smp.next1D() >= q) { done = true; endKind = 1; }
                        else T = T / q;
                    }
                }
            }
            // ---- head of the loop: path.cpp:135-222 ----
            if (!done && !(depth <= rp.maxDepth || rp.maxDepth < 0)) done = true;
            if (!done && !valid) { // camera ray missed (:136-143): environment radiance unless hidden
                if (sc.envEmitter >= 0 && !rp.hideEmitters) {
                    const DEmitter &em = sc.emitters[sc.envEmitter];
                    LiAdd = T * ((TEX && sc.envmap) ? envEvalCamera(sc, rp.diffScale, pool.pos[i], smp) : V3(em.radiance[0], em.radiance[1], em.radiance[2]));
                    liTouched = true;
                }
                done = true;
            }
            if (!done) {
                const DMaterial *mats = sc.materials;
                const int mat = its.material;
                const uint32_t btype = mats[mat].flags;
                if (its.emitter >= 0 && fresh && (!rp.hideEmitters || (flags & PF_SCATTERED))) {
                    const DEmitter &em = sc.emitters[its.emitter];
                    Spectrum le = dot(its.sh.n, -rayD) <= 0 ? Spectrum(0.0f) : V3(em.radiance[0], em.radiance[1], em.radiance[2]);
                    LiAdd = T * le; // camera ray only (fresh): the MIS term above cannot have fired
                    liTouched = true;
                }
                if ((depth >= rp.maxDepth && rp.maxDepth > 0) || (rp.strictNormals && dot(rayD, its.geoN) * cosTheta(its.wi) >= 0)) done = true;
                if (!done) {
                    V3 refN(0.0f);
                    if ((btype & (ETransmission | EBackSide)) == 0) refN = its.sh.n; // records.inl:160-164
                    // textured scenes: resolve the diffuse leaf this intersection will evaluate (twosided picks a side by wi, coating wraps
                    // its child) and look its bitmap texture up once
                    bool hasTex = false;
                    V3 texR(0.0f);
                    if (TEX) {
                        int leaf = mat;
                        if (mats[leaf].type == 5) leaf = cosTheta(its.wi) > 0 ? mats[leaf].nested : mats[leaf].nested2;
                        if (mats[leaf].type == 3) leaf = mats[leaf].nested;
                        const int tex = mats[leaf].tex;
                        if (tex >= 0) { // diffuse `reflectance`, (rough)conductor `specularReflectance`, plastic `diffuseReflectance` (the host binds nothing else)
                            const uint32_t item = sc.nItems ? pool.inst[i] : 0xFFFFFFFFu;
                            const DInstance *in = (item != 0xFFFFFFFFu && !sc.items[item].identity) ? &sc.items[item] : nullptr;
                            hasTex = true;
                            texR = shadeTexture(sc, rp, tex, prim, hit.y, hit.z, in, its, fresh, fresh ? pool.pos[i] : make_float2(0.0f, 0.0f), smp);
                        }
                    }
                    // ---- direct illumination: path.cpp:172-200 ----
                    if (btype & ESmooth) {
                        float sx, sy;
                        smp.next2D(sx, sy);
                        DirectSample ds;
                        if (sc.nEmitters > 0 && sampleEmitterDirect(sc, its.p, refN, sx, sy, ds, nullptr, TEX)) {
                            ++nShadowRef; // the reference traces (and counts, skdtree.cpp:210) the shadow ray before BSDF::eval
                            BRec bRec;
                            if (TEX) { bRec.hasTex = hasTex; bRec.texR = texR; }
                            bRec.wi = its.wi;
                            bRec.wo = its.sh.toLocal(ds.d);
                            const Spectrum bsdfVal = bsdfEval<CLS>(mats, mat, bRec);
                            if (!isZero(bsdfVal) && (!rp.strictNormals || dot(its.geoN, ds.d) * cosTheta(bRec.wo) > 0)) {
                                const float bp = bsdfPdf<CLS>(mats, mat, bRec);
                                const float weight = miWeight(ds.pdf, bp);
                                shC = T * ds.value * bsdfVal * weight;
                                shD = ds.d;
                                shMaxt = ds.dist * (1 - B2_SHADOW_EPSILON);
                                shO = its.p;
                                emitShadow = true; // the visibility test of Scene::sampleEmitterDirect (scene.cpp:838-843) is k_occluded's
                            }
                        }
                    }
                    // ---- BSDF sampling: path.cpp:206-226 ----
                    float bsdfPdfNew;
                    BRec bRec;
                    if (TEX) { bRec.hasTex = hasTex; bRec.texR = texR; }
                    bRec.wi = its.wi;
                    float sx, sy;
                    smp.next2D(sx, sy);
                    const Spectrum bsdfWeight = bsdfSample<CLS>(mats, mat, bRec, bsdfPdfNew, sx, sy, smp);
                    if (isZero(bsdfWeight)) { done = true; endKind = 2; }
                    else {
                        transmitted = (bRec.sampledType & ETransmission) != 0;
                        flags &= ~(PF_DELTA | PF_REFN_OK | PF_FRESH);
                        if (bRec.sampledType != ENull) flags |= PF_SCATTERED;
                        const V3 wo = its.sh.toWorld(bRec.wo);
                        const float woDotGeoN = dot(its.geoN, wo);
                        if (rp.strictNormals && woDotGeoN * cosTheta(bRec.wo) <= 0) { done = true; endKind = 2; }
                        else {
                            if (bRec.sampledType & EDelta) flags |= PF_DELTA;
                            if (dot(wo, refN) >= 0) flags |= PF_REFN_OK;
                            pool.ray[2 * (size_t) i] = make_float4(its.p.x, its.p.y, its.p.z, isZero(refN) ? 2.0f : dot(wo, refN));
                            pool.ray[2 * (size_t) i + 1] = make_float4(wo.x, wo.y, wo.z, bsdfPdfNew); // w: pdf of this sample (maxt = inf)
                            T = T * bsdfWeight; // :252-253 (applied early: only read again if the next ray hits)
                            eta *= bRec.eta;
                            bsdfPdfOut = bsdfPdfNew;
                        }
                    }
                }
            }
            if (smp.overflow) ++nDimOvf;
            if (done && endKind == 0) endKind = 3;
            if (rp.pathTrace && vertex >= 1 && vertex <= 8) { // diagnostics: one event byte per bounce (include/b2mts.h)
                const uint32_t pixel = pool.pix[i];
                const uint32_t sIdx = (uint32_t) (smp.index >> (2u * rp.logRes)) - (uint32_t) rp.sampleLo; // sobol look_up: index = (sample << 2m) | ...
                const size_t key = ((size_t) (pixel >> 16) * sc.cam.W + (pixel & 0xFFFFu)) * (size_t) (rp.sampleHi - rp.sampleLo) + sIdx;
                const uint32_t ev = 0x80u | (valid ? ((uint32_t) its.material & 7u) : 7u) | (emitShadow ? 8u : 0u) | (endKind << 4) | (transmitted ? 0x40u : 0u);
                rp.pathTrace[key] |= (unsigned long long) ev << (8 * (vertex - 1));
            }
            flags &= ~PF_FRESH;
            if (done) flags = (flags & ~PF_ALIVE) | PF_DONE;
            (void) bsdfPdfOut;
            pool.st[2 * (size_t) i] = make_float4(T.x, T.y, T.z, eta);
            // Li += ... (path.cpp:150,263): a fire-and-forget reduction at L2 instead of a load / add / store round trip; k_occluded adds
            // this vertex's direct illumination (path.cpp:197) with the same instruction afterwards, i.e. in the reference's order
            if (liTouched) redAddV4(&pool.st[2 * (size_t) i + 1], LiAdd.x, LiAdd.y, LiAdd.z, 0.0f);
            state = flags | ((uint32_t) depth << 8) | (smp.dim << 20);
            pool.flags[i] = state;
        }
        // finished paths: queue the slot for the next k_generate (splat + refill); shadow-ray compaction.  Warp ballot, one atomic per
        // warp and queue.  The atomics' results are NOT consumed here: the records are parked in shared memory and written out at the
        // end of the next loop iteration, when the queue positions have long arrived (ncu, round 2: 20 % of this kernel's stall samples
        // sat on the two shuffles that broadcast the freshly returned atomic).
        const bool fin = live && (state & PF_DONE);
        flushPending();
        pendA = __ballot_sync(0xffffffffu, fin); pendB = __ballot_sync(0xffffffffu, emitShadow);
        if ((tid & 31) == 0) { // (inline PTX: nvcc turns atomicAdd into elect + ATOMG + SHFL, and that shuffle would wait for the result here)
            if (pendA) baseA = atomAddPending(pool.counters + ((it & 1u) ? CTR_DONE1 : CTR_DONE0), (uint32_t) __popc(pendA));
            if (pendB) baseB = atomAddPending(pool.counters + CTR_SHADOW, (uint32_t) __popc(pendB));
        }
        if (fin) { sDoneSlot[tid] = i; ++nDone; }
        if (emitShadow) { // 48-byte shadow record: origin, direction + maxt, contribution + slot (k_occluded touches nothing else)
            sShO[tid] = make_float4(shO.x, shO.y, shO.z, 0.0f);
            sShD[tid] = make_float4(shD.x, shD.y, shD.z, shMaxt);
            sShC[tid] = make_float4(shC.x, shC.y, shC.z, __uint_as_float(i));
        }
    }
    flushPending();
#ifdef B2_STAGE_SHADE
    cpAsyncWait<0>();
#endif
    nDimOvf = warpSum(nDimOvf);
    nShadowRef = warpSum(nShadowRef);
    nDone = warpSum(nDone);
    if ((threadIdx.x & 31) == 0) {
        if (nDone) atomicAdd(pool.counters + CTR_ACTIVE, ~(unsigned long long) nDone + 1ull); // -= nDone
        if (nDimOvf) atomicAdd(pool.counters + CTR_DIMOVF, (unsigned long long) nDimOvf);
        if (nShadowRef) atomicAdd(pool.counters + CTR_SHADOWRAYS, (unsigned long long) nShadowRef);
    }
    stampEnd(rp, it, STAGE_SHADE);
}

// ------------------------------------------------------------------------------------------------
// k_occluded
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(B2_TRACE_BLOCK) k_occluded(DScene sc, DPool pool, DRender rp) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t it = (uint32_t) pool.counters[CTR_ITER] - 1u;
    stampBegin(rp, it, STAGE_OCCLUDED);
    const bool wide = sc.nodes8 != nullptr && !sc.rootCount && !sc.nItems;
    const TraceMem tm = setupTraceMem(sc, smem, wide);
    const uint32_t n = (uint32_t) pool.counters[CTR_SHADOW];
    uint32_t nClear = 0;
    if (!sc.rootCount && !sc.nItems) {
        uint32_t nv = 0, pt = 0;
        float4 cur = make_float4(0, 0, 0, 0); // contribution + slot of the ray this lane is tracing
        auto fetch = [&](uint32_t j, V3 &o, V3 &d, float &mint, float &maxt) -> int {
            const float4 ro = pool.shO[j], sd = pool.shD[j];
            cur = pool.shC[j];
            o = V3(ro.x, ro.y, ro.z); d = V3(sd.x, sd.y, sd.z);
            const V3 dRcp(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
            return clipRay<true>(sc, o, d, dRcp, B2_EPSILON, sd.w, mint, maxt) ? 2 : 1;
        };
        auto commit = [&](uint32_t, bool found, const HitRec &) {
            if (!found) { // Li += value * bsdfVal * weight (path.cpp:197): reduction at L2, no read of the slot
                redAddV4(&pool.st[2 * (size_t) __float_as_uint(cur.w) + 1], cur.x, cur.y, cur.z, 0.0f);
                ++nClear;
            }
        };
        if (wide) traverseQueue8<true, false>(sc, tm, n, pool.counters + CTR_TICKET_OCC, fetch, commit, nv, pt);
        else traverseQueue<true, false>(sc, tm, n, pool.counters + CTR_TICKET_OCC, fetch, commit, nv, pt);
    } else
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const uint32_t j = base + threadIdx.x;
        if (j < n) {
            const float4 ro = pool.shO[j], sd = pool.shD[j], scn = pool.shC[j];
            const uint32_t slot = __float_as_uint(scn.w);
            const V3 o(ro.x, ro.y, ro.z), d(sd.x, sd.y, sd.z);
            const V3 dRcp(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
            float mint, maxt;
            bool occluded = false;
            HitRec h;
            uint32_t nv = 0, pt = 0;
            uint32_t item = 0;
            if (clipRay<true>(sc, o, d, dRcp, B2_EPSILON, sd.w, mint, maxt))
                occluded = sc.nItems ? traverseTop<true, false>(sc, tm, o, d, mint, maxt, h, item, nv, pt) : traverse<true, false>(sc, tm, o, d, mint, maxt, h, nv, pt);
            if (!occluded) {
                redAddV4(&pool.st[2 * (size_t) slot + 1], scn.x, scn.y, scn.z, 0.0f);
                ++nClear;
            }
        }
    }
    nClear = warpSum(nClear);
    if ((threadIdx.x & 31) == 0 && nClear) atomicAdd(pool.counters + CTR_UNOCCLUDED, (unsigned long long) nClear);
    stampEnd(rp, it, STAGE_OCCLUDED);
}


// ------------------------------------------------------------------------------------------------
// volpath: src/integrators/path/volpath.cpp:84-366.  One launch advances every live path by one loop
// iteration (plus any chain of index-matched boundary crossings, which the reference handles with `continue`).  The
// medium work is sequential per path and data dependent (Woodcock walks draw a variable number of random numbers
// BEFORE the direction is sampled), so the rays of an iteration are cast inline by the thread that owns the path.
// ------------------------------------------------------------------------------------------------
struct VolEnv {
    const DScene &sc;
    const TraceMem &tm;
    uint32_t nRays, nShadow;
    B2_DEV VolEnv(const DScene &s, const TraceMem &t) : sc(s), tm(t), nRays(0), nShadow(0) {}
    // Scene::rayIntersect(ray, its): skdtree.cpp:112-142
    // not inlined: one copy of the traversal for the six call sites (see woodcockWalk)
    __device__ __noinline__ bool closest(const V3 &o, const V3 &d, float rayMint, float rayMaxt, HitRec &h) {
        h.t = B2_INF; h.u = 0; h.v = 0; h.prim = 0xFFFFFFFFu;
        const V3 dRcp(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
        float mint, maxt;
        uint32_t nv = 0, pt = 0;
        ++nRays;
        if (clipRay<false>(sc, o, d, dRcp, rayMint, rayMaxt, mint, maxt) && traverse<false, false>(sc, tm, o, d, mint, maxt, h, nv, pt)) return true;
        h.t = B2_INF; h.prim = 0xFFFFFFFFu;
        return false;
    }
    // ShapeKDTree::rayIntersect(ray, t, shape, n, uv): skdtree.cpp:144-204 -- closest hit, epsilon scale without the inner
    // max, unflipped face normal
    __device__ __noinline__ bool closestNormal(const V3 &o, const V3 &d, float rayMint, float rayMaxt, float &t, uint32_t &prim, V3 &n) {
        HitRec h;
        h.t = B2_INF; h.u = 0; h.v = 0; h.prim = 0xFFFFFFFFu;
        const V3 dRcp(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
        float mint, maxt;
        uint32_t nv = 0, pt = 0;
        ++nShadow;
        t = B2_INF;
        if (clipRay<true>(sc, o, d, dRcp, rayMint, rayMaxt, mint, maxt) && traverse<false, false>(sc, tm, o, d, mint, maxt, h, nv, pt)) {
            t = h.t; prim = h.prim;
            const float4 a0 = __ldg(&sc.verts[3 * (size_t) prim]), a1 = __ldg(&sc.verts[3 * (size_t) prim + 1]), a2 = __ldg(&sc.verts[3 * (size_t) prim + 2]);
            const V3 p0(a0.x, a0.y, a0.z), p1(a1.x, a1.y, a1.z), p2(a2.x, a2.y, a2.z);
            n = normalize(cross(p1 - p0, p2 - p0));
            return true;
        }
        return false;
    }
    B2_DEV int2 mediaOf(uint32_t prim) const { return sc.primMedia ? __ldg(sc.primMedia + prim) : make_int2(-1, -1); }
    B2_DEV int materialOf(uint32_t prim) const { return __float_as_int(__ldg(&sc.verts[3 * (size_t) prim].w)); }
    B2_DEV int emitterOf(uint32_t prim) const { return __float_as_int(__ldg(&sc.verts[3 * (size_t) prim + 1].w)); }
};
B2_DEV int targetMedium(const int2 &media, const V3 &geoN, const V3 &d) { return dot(d, geoN) > 0 ? media.y : media.x; } // records.inl:81-86

// Scene::evalTransmittance: scene.cpp:619-679
B2_DEV Spectrum volEvalTransmittance(VolEnv &env, const V3 &p1, bool p1OnSurface, const V3 &p2, bool p2OnSurface, int medium, int maxInteractions,
                                     PathSampler &smp) {
    const DScene &sc = env.sc;
    V3 d = p2 - p1;
    float remaining = length(d);
    d = d / remaining;
    const float lengthFactor = p2OnSurface ? (1 - B2_SHADOW_EPSILON) : 1;
    V3 o = p1;
    float rayMint = p1OnSurface ? B2_EPSILON : 0.0f, rayMaxt = remaining * lengthFactor;
    Spectrum transmittance(1.0f);
    int interactions = 0;
    while (remaining > 0) {
        float t;
        uint32_t prim = 0;
        V3 n;
        const bool surface = env.closestNormal(o, d, rayMint, rayMaxt, t, prim, n);
        if (surface && (interactions == maxInteractions || !(sc.materials[env.materialOf(prim)].flags & ENull))) return Spectrum(0.0f);
        if (medium >= 0) transmittance = transmittance * mediumTransmittance(sc.media[medium], o, d, 0.0f, fminf(t, remaining), smp);
        if (!surface || isZero(transmittance)) break;
        const int2 media = env.mediaOf(prim); // null BSDF: eval(bRec, EDiscrete) with typeMask = ENull is 1
        if (media.x >= 0 || media.y >= 0) {
            if (medium != targetMedium(media, n, -d)) return Spectrum(0.0f); // mediumInconsistencies
            medium = targetMedium(media, n, d);
        }
        if (++interactions > 100) break;
        o = o + d * t;
        remaining -= t;
        rayMaxt = remaining * lengthFactor;
        rayMint = B2_EPSILON;
    }
    return transmittance;
}

struct EmitterQuery { // DirectSamplingRecord fields pdfEmitterDirect reads (records.inl:171-179)
    V3 d, n;
    float dist;
    int emitter;
};
// volpath.cpp:368-426: first intersection into `hit` (and `rayD`-relative), attenuated emitter radiance in `value`
B2_DEV void volIntersectAndLookForEmitter(VolEnv &env, PathSampler &smp, int medium, int maxInteractions, V3 o, const V3 &d, float rayMint, HitRec &hit,
                                          EmitterQuery &eq, Spectrum &value) {
    const DScene &sc = env.sc;
    HitRec h2;
    HitRec *cur = &hit;
    Spectrum transmittance(1.0f);
    bool surface = false;
    int interactions = 0;
    while (true) {
        surface = env.closest(o, d, rayMint, B2_INF, *cur);
        if (medium >= 0) transmittance = transmittance * mediumTransmittance(sc.media[medium], o, d, 0.0f, cur->t, smp);
        if (surface) {
            const uint32_t prim = cur->prim;
            if (interactions == maxInteractions || !(sc.materials[env.materialOf(prim)].flags & ENull) || env.emitterOf(prim) >= 0) break;
        } else break;
        if (isZero(transmittance)) return;
        const int2 media = env.mediaOf(cur->prim);
        if (media.x >= 0 || media.y >= 0) {
            Isect its; // its->geoFrame.n of the full record (flipped towards the shading normal)
            fillIntersection(sc, d, cur->prim, cur->u, cur->v, its);
            medium = targetMedium(media, its.geoN, d);
        }
        o = o + d * cur->t;
        rayMint = B2_EPSILON;
        cur = &h2;
        if (++interactions > 100) return;
    }
    if (surface) {
        const int em = env.emitterOf(cur->prim);
        if (em >= 0) {
            Isect its;
            fillIntersection(sc, d, cur->prim, cur->u, cur->v, its);
            eq.n = its.sh.n; eq.d = d; eq.dist = cur->t; eq.emitter = em;
            const DEmitter &e = sc.emitters[em];
            const Spectrum le = dot(its.sh.n, -d) <= 0 ? Spectrum(0.0f) : V3(e.radiance[0], e.radiance[1], e.radiance[2]); // area.cpp:104-109
            value = transmittance * le;
        }
    } else if (sc.envEmitter >= 0) { // volpath.cpp:418-424: the ray left the scene -> environment emitter (constant.cpp:239-253)
        const DEmitter &e = sc.emitters[sc.envEmitter];
        eq.d = d; eq.n = V3(0.0f); eq.dist = 0.0f; eq.emitter = sc.envEmitter; // the solid-angle density needs d only
        value = transmittance * (sc.envmap ? envEval(*sc.envmap, sc.ewaLut, d, false, V3(0.0f), V3(0.0f)) : V3(e.radiance[0], e.radiance[1], e.radiance[2]));
    }
}
// scene.cpp:949-952; area.cpp:175-183; shape.cpp:117-126; constant.cpp:210-224
B2_DEV float volPdfEmitterDirect(const DScene &sc, const EmitterQuery &eq, const V3 &refN) {
    const DEmitter &em = sc.emitters[eq.emitter];
    if (em.nTri == 0) {
        const float pdfSA = sc.envmap ? envPdfDirection(*sc.envmap, envToLocal(*sc.envmap, eq.d)) // envmap.cpp:545-556
                                      : !isZero(refN) ? B2_INV_PI * fmaxf(0.0f, dot(eq.d, refN)) : 0.07957747154594766788f;
        return pdfSA * (em.samplingWeight * sc.emitterNormalization);
    }
    float pdfDirect = 0.0f;
    if (dot(eq.d, refN) >= 0 && dot(eq.d, eq.n) < 0) pdfDirect = em.invSurfaceArea * (eq.dist * eq.dist) / absDot(eq.d, eq.n);
    return pdfDirect * (em.samplingWeight * sc.emitterNormalization);
}

#ifndef B2_VOL_MINBLOCKS
#define B2_VOL_MINBLOCKS 5 // 256 x 5 threads per SM, <= 51 registers: measured best (the kernel is latency bound; sweep 2..8 in DESIGN.md)
#endif
__global__ void __launch_bounds__(B2_TRACE_BLOCK, B2_VOL_MINBLOCKS) k_volstep(DScene sc, DPool pool, DRender rp) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t it = (uint32_t) pool.counters[CTR_ITER] - 1u;
    stampBegin(rp, it, STAGE_SHADE);
    const TraceMem tm = setupTraceMem(sc, smem);
    const uint32_t Q = pool.capacity;
    VolEnv env(sc, tm);
    uint32_t nDimOvf = 0, nDone = 0;
    const DMaterial *mats = sc.materials;
#ifndef B2_VOL_STATIC
    // Slots are handed out per lane from a ticket counter (zeroed by k_publish): a lane that finishes its path vertex early
    // takes the next slot instead of idling until the slowest Woodcock walk of its warp is done; lanes that sit in the same
    // inner loop for different slots still issue together.
    for (;;) {
        const uint32_t i = (uint32_t) atomicAdd(pool.counters + CTR_TICKET_EXT, 1ull);
        if (i >= Q) break;
#else
    for (uint32_t base = blockIdx.x * blockDim.x; base < Q; base += gridDim.x * blockDim.x) {
        const uint32_t i = base + threadIdx.x;
#endif
        uint32_t state = i < Q ? pool.flags[i] : 0u;
        const bool live = (state & PF_ALIVE) != 0;
        if (live) {
            uint32_t flags = state & 0xFFu;
            int depth = (int) ((state >> 8) & 0xFFFu);
            const float4 ro4 = pool.ray[2 * (size_t) i], rd4 = pool.ray[2 * (size_t) i + 1], thr4 = pool.st[2 * (size_t) i];
            float4 li4 = pool.st[2 * (size_t) i + 1];
            const uint2 sm2 = pool.smp[i];
            V3 rayO(ro4.x, ro4.y, ro4.z), rayD(rd4.x, rd4.y, rd4.z);
            Spectrum T(thr4.x, thr4.y, thr4.z), Li(li4.x, li4.y, li4.z);
            float eta = thr4.w;
            PathSampler smp;
            smp.kind = rp.sampler; smp.m32 = sc.sobolNib; smp.nNib = rp.indexNibbles; smp.overflow = false; smp.cacheDim = 0xFFFFFFFFu;
            smp.index = ((uint64_t) sm2.y << 32) | sm2.x;
            smp.scramble32 = rp.sampler == 0 ? (uint32_t) rp.scramble : (uint32_t) (rp.scramble >> 32);
            int medium;
            HitRec hit;
            float rayMintCur = ro4.w, rayMaxtCur = B2_INF; // ray.mint / ray.maxt of the current path segment
            if (flags & PF_FRESH) {
                rayMaxtCur = rd4.w;
                smp.dim = state >> 20;
                medium = -1; // sensor medium: vacuum (a camera inside a medium is out of scope)
                env.closest(rayO, rayD, ro4.w, rd4.w, hit); // rRec.rayIntersect(ray), volpath.cpp:97
                if (hit.prim != 0xFFFFFFFFu) flags |= PF_ALPHA;
                flags = (flags & ~PF_FRESH) | PF_CAMRAY;
            } else {
                const uint2 v = pool.vol[i];
                medium = (int) v.x; smp.dim = v.y;
                const float4 h4 = pool.hit[i];
                hit.t = h4.x; hit.u = h4.y; hit.v = h4.z; hit.prim = __float_as_uint(h4.w);
            }
            bool done = false;
            while (true) { // one loop iteration of volpath.cpp:104-357; repeats only after an index-matched boundary (`continue`)
                if (!(depth <= rp.maxDepth || rp.maxDepth < 0)) { done = true; break; }
                const bool scattered = (flags & PF_SCATTERED) != 0;
                MediumRec mRec;
                mRec.transmittance = Spectrum(1.0f); mRec.pdfFailure = 1.0f; mRec.pdfSuccess = 1.0f;
                if (medium >= 0 && mediumSampleDistance(sc.media[medium], rayO, rayD, 0.0f, hit.t, mRec, smp)) {
                    const DMedium &med = sc.media[medium];
                    if (depth >= rp.maxDepth && rp.maxDepth != -1) { done = true; break; }
                    T = T * (mRec.sigmaS * mRec.transmittance / mRec.pdfSuccess);
                    // ---- luminaire sampling (volpath.cpp:122-151) ----
                    if (sc.nEmitters > 0) {
                        const int interactions = rp.maxDepth - depth - 1;
                        float sx, sy;
                        smp.next2D(sx, sy);
                        DirectSample ds;
                        float emPdf = 1.0f;
                        if (sampleEmitterDirect(sc, mRec.p, V3(0.0f), sx, sy, ds, &emPdf)) {
                            Spectrum value = ds.value * (volEvalTransmittance(env, mRec.p, false, ds.p, true, medium, interactions, smp) / emPdf);
                            ds.pdf *= emPdf;
                            if (!isZero(value)) {
                                const float phaseVal = phaseEval(med, -rayD, ds.d);
                                if (phaseVal != 0) Li = Li + T * value * phaseVal * miWeight(ds.pdf, phaseVal);
                            }
                        }
                    }
                    // ---- phase function sampling (volpath.cpp:153-180) ----
                    float phasePdf;
                    V3 wo;
                    const float phaseVal = phaseSample(med, -rayD, wo, phasePdf, smp);
                    if (phaseVal == 0) { done = true; break; }
                    T = T * phaseVal;
                    rayO = mRec.p; rayD = wo; rayMintCur = 0.0f; rayMaxtCur = B2_INF; flags &= ~PF_CAMRAY;
                    Spectrum value(0.0f);
                    EmitterQuery eq;
                    volIntersectAndLookForEmitter(env, smp, medium, rp.maxDepth - depth - 1, rayO, rayD, 0.0f, hit, eq, value);
                    if (!isZero(value)) Li = Li + T * value * miWeight(phasePdf, volPdfEmitterDirect(sc, eq, V3(0.0f)));
                } else {
                    if (medium >= 0) T = T * (mRec.transmittance / mRec.pdfFailure);
                    if (hit.prim == 0xFFFFFFFFu) { // volpath.cpp:190-202
                        if (sc.envEmitter >= 0 && !scattered && !rp.hideEmitters) {
                            const DEmitter &em = sc.emitters[sc.envEmitter];
                            Spectrum value = T * (!sc.envmap ? V3(em.radiance[0], em.radiance[1], em.radiance[2])
                                                             : (flags & PF_CAMRAY) ? envEvalCamera(sc, rp.diffScale, pool.pos[i], smp)
                                                                                   : envEval(*sc.envmap, sc.ewaLut, rayD, false, V3(0.0f), V3(0.0f)));
                            if (medium >= 0) value = value * mediumTransmittance(sc.media[medium], rayO, rayD, rayMintCur, rayMaxtCur, smp);
                            Li = Li + value;
                        }
                        done = true; break;
                    }
                    Isect its;
                    fillIntersection(sc, rayD, hit.prim, hit.u, hit.v, its);
                    const int mat = its.material;
                    const uint32_t btype = mats[mat].flags;
                    if (its.emitter >= 0 && !scattered /* EEmittedRadiance <=> !scattered */ && (!rp.hideEmitters || scattered)) {
                        const DEmitter &em = sc.emitters[its.emitter];
                        const Spectrum le = dot(its.sh.n, -rayD) <= 0 ? Spectrum(0.0f) : V3(em.radiance[0], em.radiance[1], em.radiance[2]);
                        Li = Li + T * le;
                    }
                    if (depth >= rp.maxDepth && rp.maxDepth != -1) { done = true; break; }
                    if ((-dot(its.geoN, rayD)) * cosTheta(its.wi) < 0 && rp.strictNormals) { done = true; break; }
                    V3 refN(0.0f);
                    if ((btype & (ETransmission | EBackSide)) == 0) refN = its.sh.n; // records.inl:160-164
                    const int2 media = env.mediaOf(hit.prim);
                    const bool transition = media.x >= 0 || media.y >= 0;
                    // ---- luminaire sampling (volpath.cpp:238-272) ----
                    if (sc.nEmitters > 0 && (btype & ESmooth)) {
                        const int interactions = rp.maxDepth - depth - 1;
                        float sx, sy;
                        smp.next2D(sx, sy);
                        DirectSample ds;
                        float emPdf = 1.0f;
                        if (sampleEmitterDirect(sc, its.p, refN, sx, sy, ds, &emPdf)) {
                            int med = medium;
                            if (transition) med = targetMedium(media, its.geoN, ds.d);
                            Spectrum value = ds.value * (volEvalTransmittance(env, its.p, true, ds.p, true, med, interactions, smp) / emPdf);
                            ds.pdf *= emPdf;
                            if (!isZero(value)) {
                                BRec bRec;
                                bRec.wi = its.wi;
                                bRec.wo = its.sh.toLocal(ds.d);
                                const Spectrum bsdfVal = bsdfEval<-1>(mats, mat, bRec);
                                if (!isZero(bsdfVal) && (!rp.strictNormals || dot(its.geoN, ds.d) * cosTheta(bRec.wo) > 0)) {
                                    const float bp = bsdfPdf<-1>(mats, mat, bRec);
                                    Li = Li + T * value * bsdfVal * miWeight(ds.pdf, bp);
                                }
                            }
                        }
                    }
                    // ---- BSDF sampling (volpath.cpp:279-300) ----
                    BRec bRec;
                    bRec.wi = its.wi;
                    float bsdfPdfNew;
                    float sx, sy;
                    smp.next2D(sx, sy);
                    const Spectrum bsdfWeight = bsdfSample<-1>(mats, mat, bRec, bsdfPdfNew, sx, sy, smp);
                    if (isZero(bsdfWeight)) { done = true; break; }
                    const V3 wo = its.sh.toWorld(bRec.wo);
                    if (dot(its.geoN, wo) * cosTheta(bRec.wo) <= 0 && rp.strictNormals) { done = true; break; }
                    rayO = its.p; rayD = wo; rayMintCur = B2_EPSILON; rayMaxtCur = B2_INF; flags &= ~PF_CAMRAY;
                    T = T * bsdfWeight;
                    eta *= bRec.eta;
                    if (transition) medium = targetMedium(media, its.geoN, rayD);
                    if (bRec.sampledType == ENull) { // index-matched boundary: volpath.cpp:302-311
                        env.closest(rayO, rayD, B2_EPSILON, B2_INF, hit);
                        depth++;
                        continue;
                    }
                    Spectrum value(0.0f);
                    EmitterQuery eq;
                    volIntersectAndLookForEmitter(env, smp, medium, rp.maxDepth - depth - 1, rayO, rayD, B2_EPSILON, hit, eq, value);
                    if (!isZero(value)) {
                        const float emitterPdf = !(bRec.sampledType & EDelta) ? volPdfEmitterDirect(sc, eq, refN) : 0.0f;
                        Li = Li + T * value * miWeight(bsdfPdfNew, emitterPdf);
                    }
                }
                if (depth++ >= rp.rrDepth) { // volpath.cpp:345-354
                    const float q = fminf(maxComp(T) * eta * eta, 0.95f);
                    if (smp.next1D() >= q) { done = true; break; }
                    T = T / q;
                }
                flags |= PF_SCATTERED;
                // not in the reference (it aborts the render, sobol.cpp:223-225): a path that ran past the Sobol' table ends here
                if (smp.overflow) done = true;
                break;
            }
            if (smp.overflow) ++nDimOvf;
            if (done) flags = (flags & ~PF_ALIVE) | PF_DONE;
            else {
                pool.ray[2 * (size_t) i] = make_float4(rayO.x, rayO.y, rayO.z, rayMintCur);
                pool.ray[2 * (size_t) i + 1] = make_float4(rayD.x, rayD.y, rayD.z, 0.0f);
                pool.hit[i] = make_float4(hit.t, hit.u, hit.v, __uint_as_float(hit.prim));
                pool.st[2 * (size_t) i] = make_float4(T.x, T.y, T.z, eta);
                pool.vol[i] = make_uint2((uint32_t) medium, smp.dim);
            }
            pool.st[2 * (size_t) i + 1] = make_float4(Li.x, Li.y, Li.z, 0.0f);
            state = flags | ((uint32_t) min(depth, 0xFFF) << 8);
            pool.flags[i] = state;
        }
        const bool fin = live && (state & PF_DONE);
        const uint32_t dq = warpAppend(fin, pool.counters + ((it & 1u) ? CTR_DONE1 : CTR_DONE0));
        if (fin) {
            pool.doneQueue[(size_t) (it & 1u) * Q + dq] = i;
            ++nDone;
        }
    }
    nDimOvf = warpSum(nDimOvf);
    nDone = warpSum(nDone);
    const uint32_t nRays = warpSum(env.nRays), nShadow = warpSum(env.nShadow);
    if ((threadIdx.x & 31) == 0) {
        if (nDone) atomicAdd(pool.counters + CTR_ACTIVE, ~(unsigned long long) nDone + 1ull);
        if (nDimOvf) atomicAdd(pool.counters + CTR_DIMOVF, (unsigned long long) nDimOvf);
        if (nRays) atomicAdd(pool.counters + CTR_RAYS, (unsigned long long) nRays);
        if (nShadow) atomicAdd(pool.counters + CTR_SHADOWRAYS, (unsigned long long) nShadow);
    }
    stampEnd(rp, it, STAGE_SHADE);
}


// ------------------------------------------------------------------------------------------------
// k_volstep_lockstep: the same loop iteration as k_volstep, arranged so that the 32 paths of a warp walk through it TOGETHER.
// The medium-scattering and the surface branch of volpath.cpp:108-343 only differ in how they evaluate the scattering function;
// their three expensive parts -- the distance-sampling walk, the attenuated shadow connection (scene.cpp:619-679) and the emitter
// look-up after sampling a direction (volpath.cpp:368-426) -- are hoisted out of the branches into ONE call site each, and slots are
// assigned statically, so every lane of a warp enters the same Woodcock loop at the same time (ncu: k_volstep ran 2.3 of 32 lanes per
// instruction).  Same draws in the same order per path as k_volstep and the oracle.
// ------------------------------------------------------------------------------------------------
#ifndef B2_VOLLS_MINBLOCKS
#define B2_VOLLS_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(B2_TRACE_BLOCK, B2_VOLLS_MINBLOCKS) k_volstep_lockstep(DScene sc, DPool pool, DRender rp) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t it = (uint32_t) pool.counters[CTR_ITER] - 1u;
    stampBegin(rp, it, STAGE_SHADE);
    const TraceMem tm = setupTraceMem(sc, smem);
    const uint32_t Q = pool.capacity;
    VolEnv env(sc, tm);
    uint32_t nDimOvf = 0, nDone = 0;
    const DMaterial *mats = sc.materials;
    __shared__ uint32_t sPerm[B2_TRACE_BLOCK];
    __shared__ uint32_t sCnt[3][B2_TRACE_BLOCK / 32];
    for (uint32_t base = blockIdx.x * blockDim.x; base < Q; base += gridDim.x * blockDim.x) {
        // The CTA partitions its 256 slots by what the next iteration will do -- paths inside a medium (long Woodcock walks), other live
        // paths, empty slots -- so that the lanes of a warp have similar work (stable partition by warp ballots + an 8-entry scan).
        uint32_t i;
        {
            const uint32_t j = base + threadIdx.x;
            const uint32_t stj = j < Q ? pool.flags[j] : 0u;
            const bool alivej = (stj & PF_ALIVE) != 0;
            const bool inside = alivej && !(stj & PF_FRESH) && (int) pool.vol[j].x >= 0;
            const int key = inside ? 0 : (alivej ? 1 : 2);
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            const unsigned m0 = __ballot_sync(0xffffffffu, key == 0), m1 = __ballot_sync(0xffffffffu, key == 1), m2 = __ballot_sync(0xffffffffu, key == 2);
            if (lane == 0) { sCnt[0][warp] = __popc(m0); sCnt[1][warp] = __popc(m1); sCnt[2][warp] = __popc(m2); }
            __syncthreads();
            uint32_t before = 0; // slots of smaller keys in the CTA + slots of my key in earlier warps
            const int nW = blockDim.x >> 5;
            for (int k = 0; k < key; ++k)
                for (int w = 0; w < nW; ++w) before += sCnt[k][w];
            for (int w = 0; w < warp; ++w) before += sCnt[key][w];
            const unsigned mine = key == 0 ? m0 : (key == 1 ? m1 : m2);
            sPerm[before + __popc(mine & ((1u << lane) - 1u))] = j;
            __syncthreads();
            i = sPerm[threadIdx.x];
            __syncthreads();
        }
        uint32_t state = i < Q ? pool.flags[i] : 0u;
        const bool live = (state & PF_ALIVE) != 0;
        uint32_t flags = state & 0xFFu;
        int depth = (int) ((state >> 8) & 0xFFFu);
        V3 rayO(0.0f), rayD(0.0f, 0.0f, 1.0f);
        Spectrum T(0.0f), Li(0.0f);
        float eta = 1.0f;
        PathSampler smp;
        smp.kind = rp.sampler; smp.m32 = sc.sobolNib; smp.nNib = rp.indexNibbles; smp.overflow = false; smp.cacheDim = 0xFFFFFFFFu;
        smp.index = 0; smp.dim = 0;
        smp.scramble32 = rp.sampler == 0 ? (uint32_t) rp.scramble : (uint32_t) (rp.scramble >> 32);
        int medium = -1;
        HitRec hit;
        hit.t = B2_INF; hit.u = 0; hit.v = 0; hit.prim = 0xFFFFFFFFu;
        float rayMintCur = 0.0f, rayMaxtCur = B2_INF;
        bool fresh = false;
        if (live) {
            const float4 ro4 = pool.ray[2 * (size_t) i], rd4 = pool.ray[2 * (size_t) i + 1], thr4 = pool.st[2 * (size_t) i], li4 = pool.st[2 * (size_t) i + 1];
            const uint2 sm2 = pool.smp[i];
            rayO = V3(ro4.x, ro4.y, ro4.z); rayD = V3(rd4.x, rd4.y, rd4.z);
            T = Spectrum(thr4.x, thr4.y, thr4.z); Li = Spectrum(li4.x, li4.y, li4.z);
            eta = thr4.w;
            smp.index = ((uint64_t) sm2.y << 32) | sm2.x;
            rayMintCur = ro4.w;
            fresh = (flags & PF_FRESH) != 0;
            if (fresh) { rayMaxtCur = rd4.w; smp.dim = state >> 20; }
            else {
                const uint2 v = pool.vol[i];
                medium = (int) v.x; smp.dim = v.y;
                const float4 h4 = pool.hit[i];
                hit.t = h4.x; hit.u = h4.y; hit.v = h4.z; hit.prim = __float_as_uint(h4.w);
            }
        }
        // rRec.rayIntersect(ray) of the camera ray (volpath.cpp:97): one call site for the warp
        if (fresh) {
            env.closest(rayO, rayD, rayMintCur, rayMaxtCur, hit);
            if (hit.prim != 0xFFFFFFFFu) flags |= PF_ALPHA;
            flags = (flags & ~PF_FRESH) | PF_CAMRAY;
        }
        // One pass per launch: a lane that crosses an index-matched boundary (the reference `continue`s, volpath.cpp:302-311) keeps its
        // state and takes the repeated loop iteration in the NEXT launch -- repeating it here would leave the rest of the warp waiting.
        bool done = false, busy = live;
        {
            // ---- A: loop condition + distance sampling (one walk for the whole warp) ----
            bool run = busy;
            if (run && !(depth <= rp.maxDepth || rp.maxDepth < 0)) { done = true; busy = false; run = false; }
            const bool scattered = (flags & PF_SCATTERED) != 0;
            MediumRec mRec;
            mRec.t = 0; mRec.p = V3(0.0f); mRec.sigmaS = Spectrum(0.0f); mRec.transmittance = Spectrum(1.0f); mRec.pdfFailure = 1.0f; mRec.pdfSuccess = 1.0f;
            bool mediumEvent = false;
            if (run && medium >= 0) mediumEvent = mediumSampleDistance(sc.media[medium], rayO, rayD, 0.0f, hit.t, mRec, smp);
            // ---- B: branch-specific preparation of the direct-illumination query ----
            Isect its;
            its.material = 0; its.emitter = -1;
            V3 neeRef(0.0f), refN(0.0f);
            bool doNee = false, neeOnSurface = false;
            int2 media = make_int2(-1, -1);
            uint32_t btype = 0;
            if (run) {
                if (mediumEvent) {
                    if (depth >= rp.maxDepth && rp.maxDepth != -1) { done = true; busy = false; run = false; }
                    else {
                        T = T * (mRec.sigmaS * mRec.transmittance / mRec.pdfSuccess);
                        neeRef = mRec.p; doNee = sc.nEmitters > 0;
                    }
                } else {
                    if (medium >= 0) T = T * (mRec.transmittance / mRec.pdfFailure);
                    if (hit.prim == 0xFFFFFFFFu) { // volpath.cpp:190-202
                        if (sc.envEmitter >= 0 && !scattered && !rp.hideEmitters) {
                            const DEmitter &em = sc.emitters[sc.envEmitter];
                            Spectrum value = T * (!sc.envmap ? V3(em.radiance[0], em.radiance[1], em.radiance[2])
                                                             : (flags & PF_CAMRAY) ? envEvalCamera(sc, rp.diffScale, pool.pos[i], smp)
                                                                                   : envEval(*sc.envmap, sc.ewaLut, rayD, false, V3(0.0f), V3(0.0f)));
                            if (medium >= 0) value = value * mediumTransmittance(sc.media[medium], rayO, rayD, rayMintCur, rayMaxtCur, smp);
                            Li = Li + value;
                        }
                        done = true; busy = false; run = false;
                    } else {
                        fillIntersection(sc, rayD, hit.prim, hit.u, hit.v, its);
                        btype = mats[its.material].flags;
                        if (its.emitter >= 0 && !scattered && (!rp.hideEmitters || scattered)) {
                            const DEmitter &em = sc.emitters[its.emitter];
                            const Spectrum le = dot(its.sh.n, -rayD) <= 0 ? Spectrum(0.0f) : V3(em.radiance[0], em.radiance[1], em.radiance[2]);
                            Li = Li + T * le;
                        }
                        if ((depth >= rp.maxDepth && rp.maxDepth != -1) || ((-dot(its.geoN, rayD)) * cosTheta(its.wi) < 0 && rp.strictNormals)) { done = true; busy = false; run = false; }
                        else {
                            if ((btype & (ETransmission | EBackSide)) == 0) refN = its.sh.n; // records.inl:160-164
                            media = env.mediaOf(hit.prim);
                            neeRef = its.p; neeOnSurface = true;
                            doNee = sc.nEmitters > 0 && (btype & ESmooth);
                        }
                    }
                }
            }
            const bool transition = media.x >= 0 || media.y >= 0;
            // ---- C: luminaire sampling, shared by both branches (volpath.cpp:122-151, 238-272) ----
            {
                DirectSample ds;
                float emPdf = 1.0f;
                bool ok = false;
                if (run && doNee) {
                    float sx, sy;
                    smp.next2D(sx, sy);
                    ok = sampleEmitterDirect(sc, neeRef, refN, sx, sy, ds, &emPdf);
                }
                int med = medium;
                if (ok && neeOnSurface && transition) med = targetMedium(media, its.geoN, ds.d);
                Spectrum tr(0.0f);
                if (ok) tr = volEvalTransmittance(env, neeRef, neeOnSurface, ds.p, true, med, rp.maxDepth - depth - 1, smp);
                if (ok) {
                    const Spectrum value = ds.value * (tr / emPdf);
                    ds.pdf *= emPdf;
                    if (!isZero(value)) {
                        if (mediumEvent) {
                            const float phaseVal = phaseEval(sc.media[medium], -rayD, ds.d);
                            if (phaseVal != 0) Li = Li + T * value * phaseVal * miWeight(ds.pdf, phaseVal);
                        } else {
                            BRec bRec;
                            bRec.wi = its.wi;
                            bRec.wo = its.sh.toLocal(ds.d);
                            const Spectrum bsdfVal = bsdfEval<-1>(mats, its.material, bRec);
                            if (!isZero(bsdfVal) && (!rp.strictNormals || dot(its.geoN, ds.d) * cosTheta(bRec.wo) > 0))
                                Li = Li + T * value * bsdfVal * miWeight(ds.pdf, bsdfPdf<-1>(mats, its.material, bRec));
                        }
                    }
                }
            }
            // ---- D: sample the next direction (phase function or BSDF) ----
            float scatterPdf = 0.0f;
            bool isNull = false, isDelta = false;
            float lookMint = 0.0f;
            if (run) {
                if (mediumEvent) {
                    V3 wo;
                    const float phaseVal = phaseSample(sc.media[medium], -rayD, wo, scatterPdf, smp);
                    if (phaseVal == 0) { done = true; busy = false; run = false; }
                    else {
                        T = T * phaseVal;
                        rayO = mRec.p; rayD = wo; rayMintCur = 0.0f; rayMaxtCur = B2_INF; flags &= ~PF_CAMRAY;
                        refN = V3(0.0f);
                    }
                } else {
                    BRec bRec;
                    bRec.wi = its.wi;
                    float sx, sy;
                    smp.next2D(sx, sy);
                    const Spectrum bsdfWeight = bsdfSample<-1>(mats, its.material, bRec, scatterPdf, sx, sy, smp);
                    const V3 wo = its.sh.toWorld(bRec.wo);
                    if (isZero(bsdfWeight) || (dot(its.geoN, wo) * cosTheta(bRec.wo) <= 0 && rp.strictNormals)) { done = true; busy = false; run = false; }
                    else {
                        rayO = its.p; rayD = wo; rayMintCur = B2_EPSILON; rayMaxtCur = B2_INF; flags &= ~PF_CAMRAY;
                        T = T * bsdfWeight;
                        eta *= bRec.eta;
                        if (transition) medium = targetMedium(media, its.geoN, rayD);
                        isNull = bRec.sampledType == ENull;
                        isDelta = (bRec.sampledType & EDelta) != 0;
                        lookMint = B2_EPSILON;
                    }
                }
            }
            // ---- E: index-matched boundary (volpath.cpp:302-311): plain intersection, no Russian roulette, repeat the iteration ----
            if (run && isNull) {
                env.closest(rayO, rayD, B2_EPSILON, B2_INF, hit);
                depth++;
                run = false; busy = false; // the next loop iteration of this path runs in the next launch
            }
            // ---- F: first intersection along the new ray + attenuated emitter behind index-matched boundaries (one call site) ----
            Spectrum value(0.0f);
            EmitterQuery eq;
            eq.emitter = -1;
            if (run) volIntersectAndLookForEmitter(env, smp, medium, rp.maxDepth - depth - 1, rayO, rayD, lookMint, hit, eq, value);
            // ---- G: MIS with the emitter's density, Russian roulette ----
            if (run) {
                if (!isZero(value)) {
                    const float emitterPdf = isDelta ? 0.0f : volPdfEmitterDirect(sc, eq, refN);
                    Li = Li + T * value * miWeight(scatterPdf, emitterPdf);
                }
                if (depth++ >= rp.rrDepth) { // volpath.cpp:345-354
                    const float q = fminf(maxComp(T) * eta * eta, 0.95f);
                    if (smp.next1D() >= q) done = true;
                    else T = T / q;
                }
                if (!done) {
                    flags |= PF_SCATTERED;
                    if (smp.overflow) done = true; // a path that ran past the Sobol' table ends here (see k_volstep)
                }
                busy = false;
            }
        }
        if (live) {
            if (smp.overflow) ++nDimOvf;
            if (done) flags = (flags & ~PF_ALIVE) | PF_DONE;
            else {
                pool.ray[2 * (size_t) i] = make_float4(rayO.x, rayO.y, rayO.z, rayMintCur);
                pool.ray[2 * (size_t) i + 1] = make_float4(rayD.x, rayD.y, rayD.z, 0.0f);
                pool.hit[i] = make_float4(hit.t, hit.u, hit.v, __uint_as_float(hit.prim));
                pool.st[2 * (size_t) i] = make_float4(T.x, T.y, T.z, eta);
                pool.vol[i] = make_uint2((uint32_t) medium, smp.dim);
            }
            pool.st[2 * (size_t) i + 1] = make_float4(Li.x, Li.y, Li.z, 0.0f);
            state = flags | ((uint32_t) min(depth, 0xFFF) << 8);
            pool.flags[i] = state;
        }
        const bool fin = live && (state & PF_DONE);
        const uint32_t dq = warpAppend(fin, pool.counters + ((it & 1u) ? CTR_DONE1 : CTR_DONE0));
        if (fin) {
            pool.doneQueue[(size_t) (it & 1u) * Q + dq] = i;
            ++nDone;
        }
    }
    nDimOvf = warpSum(nDimOvf);
    nDone = warpSum(nDone);
    const uint32_t nRays = warpSum(env.nRays), nShadow = warpSum(env.nShadow);
    if ((threadIdx.x & 31) == 0) {
        if (nDone) atomicAdd(pool.counters + CTR_ACTIVE, ~(unsigned long long) nDone + 1ull);
        if (nDimOvf) atomicAdd(pool.counters + CTR_DIMOVF, (unsigned long long) nDimOvf);
        if (nRays) atomicAdd(pool.counters + CTR_RAYS, (unsigned long long) nRays);
        if (nShadow) atomicAdd(pool.counters + CTR_SHADOWRAYS, (unsigned long long) nShadow);
    }
    stampEnd(rp, it, STAGE_SHADE);
}

// medium component probes (b2_medium_probe): what = 0 transmittance (in: 8 floats/ray -> 3), 1 sampleDistance (-> 12),
// 2 density lookup (in: 3 floats -> 1), 3 phase sample (in: wi xyz + 2 samples -> 5); counter stream keyed like the oracle probe
__global__ void k_medium_probe(DScene sc, int medium, int what, uint64_t n, const float *in, uint64_t seed, float *out) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const DMedium &m = sc.media[medium];
        PathSampler smp;
        smp.kind = 2; smp.m32 = nullptr; smp.nNib = 0; smp.overflow = false; smp.cacheDim = 0xFFFFFFFFu; smp.dim = 0;
        smp.scramble32 = (uint32_t) (seed >> 32);
        const uint32_t px = (uint32_t) (i & 0xFFFF), py = (uint32_t) (i >> 16);
        smp.index = ((py * 65536u + px) * 1u + 0u) ^ (uint32_t) seed;
        if (what == 0 || what == 1) {
            const float *r = in + 8 * i;
            const V3 o(r[0], r[1], r[2]), d(r[4], r[5], r[6]);
            if (what == 0) {
                const Spectrum t = mediumTransmittance(m, o, d, r[3], r[7], smp);
                out[3 * i] = t.x; out[3 * i + 1] = t.y; out[3 * i + 2] = t.z;
            } else {
                MediumRec mRec;
                mRec.t = 0; mRec.p = V3(0.0f); mRec.sigmaS = Spectrum(0.0f); mRec.transmittance = Spectrum(1.0f); mRec.pdfFailure = 1; mRec.pdfSuccess = 1;
                const bool ok = mediumSampleDistance(m, o, d, r[3], r[7], mRec, smp);
                float *q = out + 12 * i;
                q[0] = ok ? 1.0f : 0.0f; q[1] = mRec.t; q[2] = mRec.sigmaS.x; q[3] = mRec.sigmaS.y; q[4] = mRec.sigmaS.z;
                q[5] = mRec.transmittance.x; q[6] = mRec.transmittance.y; q[7] = mRec.transmittance.z; q[8] = mRec.pdfSuccess; q[9] = mRec.pdfFailure;
                q[10] = 0; q[11] = 0;
            }
        } else if (what == 2) {
            out[i] = lookupDensity(m, V3(in[3 * i], in[3 * i + 1], in[3 * i + 2]));
        } else {
            const float *r = in + 5 * i;
            smp.kind = 4; // two-value replay
            smp.replayA = r[3]; smp.replayB = r[4];
            V3 wo;
            float pdf;
            const V3 wi(r[0], r[1], r[2]);
            phaseSample(m, wi, wo, pdf, smp);
            float *q = out + 5 * i;
            q[0] = wo.x; q[1] = wo.y; q[2] = wo.z; q[3] = pdf; q[4] = phaseEval(m, wi, wo);
        }
    }
}

// film pack: (float4 (r, g, b, weight * alpha), float weight * (1 - alpha)) planes -> interleaved H*W*5 (hdrfilm.cpp:351-356 ESpectrumAlphaWeight)
__global__ void k_film_pack(const float4 *rgba, const float *w, float *out, size_t n) {
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        const float4 v = rgba[i]; // (r, g, b, sum of weight * alpha); w[i] = sum of weight * (1 - alpha)
        float *o = out + 5 * i;
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; o[4] = v.w + w[i];
    }
}

// ------------------------------------------------------------------------------------------------
// Ray binning (counting sort) in front of the wide-tree traversal: rays that enter the scene box in the same cell of an 8 x 8 x 8
// grid AND point into the same cell of a 16 x 16 octahedral direction map get consecutive tickets, so the lanes of a warp walk
// similar nodes (fewer idle lanes, the node and triangle lines they share are fetched once).  Three hand-written passes: key +
// histogram (one global atomic per ray over 131 072 bins), exclusive scan (one CTA), scatter (one atomic per ray).  The order inside a
// bin is arbitrary; hits do not depend on it.
// ------------------------------------------------------------------------------------------------
B2_DEV uint32_t spread3(uint32_t v) { // 3 bits -> every third bit
    return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4);
}
B2_DEV uint32_t rayBinKey(const DScene &sc, const V3 &o, const V3 &d) {
    const V3 dRcp(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    float t0, t1;
    V3 p = o;
    if (sceneBoxIntersect(sc, o, d, dRcp, t0, t1) && t0 > 0) p = o + d * t0; // entry point of a ray that starts outside the scene box
    const float ex = sc.aabbMax[0] - sc.aabbMin[0], ey = sc.aabbMax[1] - sc.aabbMin[1], ez = sc.aabbMax[2] - sc.aabbMin[2];
    const uint32_t cx = (uint32_t) fminf(fmaxf((p.x - sc.aabbMin[0]) / ex * 8.0f, 0.0f), 7.0f);
    const uint32_t cy = (uint32_t) fminf(fmaxf((p.y - sc.aabbMin[1]) / ey * 8.0f, 0.0f), 7.0f);
    const uint32_t cz = (uint32_t) fminf(fmaxf((p.z - sc.aabbMin[2]) / ez * 8.0f, 0.0f), 7.0f);
    // octahedral map of the direction
    const float inv = 1.0f / (fabsf(d.x) + fabsf(d.y) + fabsf(d.z));
    float u = d.x * inv, v = d.y * inv;
    if (d.z < 0) { const float uu = (1.0f - fabsf(v)) * (u >= 0 ? 1.0f : -1.0f), vv = (1.0f - fabsf(u)) * (v >= 0 ? 1.0f : -1.0f); u = uu; v = vv; }
    const uint32_t du = (uint32_t) fminf(fmaxf((u * 0.5f + 0.5f) * 16.0f, 0.0f), 15.0f), dv = (uint32_t) fminf(fmaxf((v * 0.5f + 0.5f) * 16.0f, 0.0f), 15.0f);
    return ((spread3(cx) | (spread3(cy) << 1) | (spread3(cz) << 2)) << 8) | (du << 4) | dv;
}
// rays: the ray array of b2_trace_device, or null = the path pool (live slots only: a dead slot gets no ticket)
__global__ void k_bin_count(DScene sc, DPool pool, const float4 *rays, uint32_t n, uint32_t *keys, uint32_t *hist) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 ro, rd;
        if (rays) { ro = rays[2 * (size_t) i]; rd = rays[2 * (size_t) i + 1]; }
        else {
            if (!(pool.flags[i] & PF_ALIVE)) { keys[i] = 0xFFFFFFFFu; continue; }
            ro = pool.ray[2 * (size_t) i]; rd = pool.ray[2 * (size_t) i + 1];
        }
        const uint32_t k = rayBinKey(sc, V3(ro.x, ro.y, ro.z), V3(rd.x, rd.y, rd.z));
        keys[i] = k;
        atomicAdd(hist + k, 1u);
    }
}
// one CTA of 1024 threads: exclusive prefix sum over the B2_NBINS counters in place; hist[B2_NBINS] = total
__global__ void __launch_bounds__(1024) k_bin_scan(uint32_t *hist) {
    __shared__ uint32_t sSum[1024];
    const uint32_t per = B2_NBINS / 1024u, base = threadIdx.x * per;
    uint32_t acc = 0;
    for (uint32_t k = 0; k < per; ++k) acc += hist[base + k];
    sSum[threadIdx.x] = acc;
    __syncthreads();
    for (uint32_t off = 1; off < 1024u; off <<= 1) { // Hillis-Steele inclusive scan of the 1024 partial sums
        const uint32_t v = threadIdx.x >= off ? sSum[threadIdx.x - off] : 0u;
        __syncthreads();
        sSum[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = threadIdx.x ? sSum[threadIdx.x - 1] : 0u;
    for (uint32_t k = 0; k < per; ++k) { const uint32_t c = hist[base + k]; hist[base + k] = run; run += c; }
    if (threadIdx.x == 1023) hist[B2_NBINS] = run;
}
__global__ void k_bin_scatter(uint32_t n, const uint32_t *keys, uint32_t *cursor, uint32_t *order) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t k = keys[i];
        if (k != 0xFFFFFFFFu) order[atomicAdd(cursor + k, 1u)] = i;
    }
}

// ------------------------------------------------------------------------------------------------
// component kernels (b2_trace / b2_bsdf_* / b2_sample_emitter_direct / b2_camera_rays / b2_sampler_stream / b2_splat)
// ------------------------------------------------------------------------------------------------
template <bool SHADOW, bool COUNT> __global__ void __launch_bounds__(B2_TRACE_BLOCK) k_trace_rays(DScene sc, const float4 *rays, float4 *out, uint64_t n,
                                                                                                   unsigned long long *counters, const uint32_t *order) {
    extern __shared__ __align__(128) unsigned char smem[];
    const bool wide = sc.nodes8 != nullptr && !sc.rootCount && !sc.nItems;
    const TraceMem tm = setupTraceMem(sc, smem, wide);
    uint32_t nv = 0, pt = 0;
    if (!sc.rootCount && !sc.nItems) {
        auto fetch = [&](uint32_t t, V3 &o, V3 &d, float &mint, float &maxt) -> int {
            const uint32_t i = order ? order[t] : t; // binned tickets (k_bin_*): neighbouring lanes get similar rays
            const float4 ro = rays[2 * (size_t) i], rd = rays[2 * (size_t) i + 1];
            o = V3(ro.x, ro.y, ro.z); d = V3(rd.x, rd.y, rd.z);
            const V3 dRcp(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
            return clipRay<SHADOW>(sc, o, d, dRcp, ro.w, rd.w, mint, maxt) ? 2 : 1;
        };
        auto commit = [&](uint32_t t, bool found, const HitRec &h) {
            const uint32_t i = order ? order[t] : t;
            if (SHADOW) out[i] = make_float4(0, 0, 0, __uint_as_float(found ? 1u : 0u));
            else out[i] = found ? make_float4(h.t, h.u, h.v, __uint_as_float(h.prim)) : make_float4(B2_INF, 0.0f, 0.0f, __uint_as_float(0xFFFFFFFFu));
        };
        if (wide) traverseQueue8<SHADOW, COUNT>(sc, tm, (uint32_t) n, counters + CTR_TICKET_EXT, fetch, commit, nv, pt);
        else traverseQueue<SHADOW, COUNT>(sc, tm, (uint32_t) n, counters + CTR_TICKET_EXT, fetch, commit, nv, pt);
    } else
    for (uint64_t base = blockIdx.x * (uint64_t) blockDim.x; base < n; base += (uint64_t) gridDim.x * blockDim.x) {
        const uint64_t i = base + threadIdx.x;
        if (i < n) {
            const float4 ro = rays[2 * i], rd = rays[2 * i + 1];
            const V3 o(ro.x, ro.y, ro.z), d(rd.x, rd.y, rd.z);
            const V3 dRcp(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
            HitRec h;
            h.t = B2_INF; h.u = 0; h.v = 0; h.prim = 0xFFFFFFFFu;
            float mint, maxt;
            bool found = false;
            uint32_t item = 0;
            if (clipRay<SHADOW>(sc, o, d, dRcp, ro.w, rd.w, mint, maxt))
                found = sc.nItems ? traverseTop<SHADOW, COUNT>(sc, tm, o, d, mint, maxt, h, item, nv, pt) : traverse<SHADOW, COUNT>(sc, tm, o, d, mint, maxt, h, nv, pt);
            if (SHADOW) out[i] = make_float4(0, 0, 0, __uint_as_float(found ? 1u : 0u));
            else {
                if (!found) { h.t = B2_INF; h.u = 0; h.v = 0; h.prim = 0xFFFFFFFFu; }
                out[i] = make_float4(h.t, h.u, h.v, __uint_as_float(h.prim));
            }
        }
    }
    if (COUNT) {
        nv = warpSum(nv); pt = warpSum(pt);
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(counters + CTR_NODEVIS, (unsigned long long) nv);
            atomicAdd(counters + CTR_PRIMTESTS, (unsigned long long) pt);
        }
    }
}

__global__ void k_bsdf_eval(DScene sc, int mat, uint64_t n, const float *wi, const float *wo, float *rgb, float *pdf) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        BRec r;
        r.wi = V3(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]);
        r.wo = V3(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]);
        const Spectrum f = bsdfEval<-1>(sc.materials, mat, r);
        rgb[3 * i] = f.x; rgb[3 * i + 1] = f.y; rgb[3 * i + 2] = f.z;
        pdf[i] = bsdfPdf<-1>(sc.materials, mat, r);
    }
}
__global__ void k_bsdf_sample(DScene sc, int mat, uint64_t n, const float *wi, const float *samples, float *out) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        // replay sampler (kind 3): next1D() returns the third sample component (test_chisquare.cpp:58-92 FakeSampler)
        PathSampler smp;
        smp.kind = 3; smp.m32 = nullptr; smp.nNib = 8; smp.overflow = false; smp.cacheDim = 0xFFFFFFFFu; smp.dim = 0; smp.index = 0;
        smp.scramble32 = __float_as_uint(samples[3 * i + 2]);
        BRec r;
        r.wi = V3(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]);
        r.wo = V3(0.0f); r.eta = 1.0f; r.sampledType = 0;
        float pdf = 0;
        const Spectrum w = bsdfSample<-1>(sc.materials, mat, r, pdf, samples[3 * i], samples[3 * i + 1], smp);
        float *o = out + 10 * i;
        o[0] = r.wo.x; o[1] = r.wo.y; o[2] = r.wo.z; o[3] = w.x; o[4] = w.y; o[5] = w.z;
        o[6] = isZero(w) ? 0.0f : pdf; o[7] = (float) r.sampledType; o[8] = r.eta; o[9] = 0;
    }
}
__global__ void k_emitter_direct(DScene sc, uint64_t n, const float *ref, const float *samples, float *out) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        DirectSample ds;
        const V3 r(ref[6 * i], ref[6 * i + 1], ref[6 * i + 2]), rn(ref[6 * i + 3], ref[6 * i + 4], ref[6 * i + 5]);
        const bool ok = sampleEmitterDirect(sc, r, rn, samples[2 * i], samples[2 * i + 1], ds);
        float *o = out + 12 * i;
        o[0] = ds.d.x; o[1] = ds.d.y; o[2] = ds.d.z; o[3] = ds.dist; o[4] = ok ? ds.pdf : 0.0f;
        o[5] = ds.value.x; o[6] = ds.value.y; o[7] = ds.value.z; o[8] = ok ? 1.0f : 0.0f; o[9] = ds.p.x; o[10] = ds.p.y; o[11] = ds.p.z;
    }
}
// Texture probes (parity tests).  what = 0: Texture2D::eval of texture `tex` (in n x 6: u, v, dudx, dudy, dvdx, dvdy; hasPartials selects the
// filtered look-up) -> out n x 3.  what = 1: uv and uv partials of camera-ray hits (in n x 6: film position, t, barycentric u, v, prim bits;
// world triangles only) -> out n x 6: u, v, dudx, dudy, dvdx, dvdy
__global__ void k_texture_probe(DScene sc, int what, int tex, int hasPartials, float diffScale, uint64_t n, const float *in, float *out) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const float *r = in + 6 * i;
        if (what == 0) {
            TexCoord tc;
            tc.u = r[0]; tc.v = r[1]; tc.hasUVPartials = hasPartials != 0;
            tc.dudx = r[2]; tc.dudy = r[3]; tc.dvdx = r[4]; tc.dvdy = r[5];
            const Spectrum v = texEval(sc.textures[tex], sc.ewaLut, tc);
            out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z;
        } else {
            V3 o, d;
            float mint, maxt;
            cameraRay(sc.cam, r[0], r[1], o, d, mint, maxt);
            Isect its;
            const uint32_t prim = __float_as_uint(r[5]);
            fillIntersection(sc, d, prim, r[3], r[4], its);
            PathSampler smp;
            smp.kind = 2; smp.index = 0; smp.dim = 0; smp.scramble32 = 0; smp.overflow = false; smp.cacheDim = 0xFFFFFFFFu;
            const TexCoord tc = surfaceTexCoord(sc, diffScale, prim, r[3], r[4], nullptr, its.p, its.geoN, true, make_float2(r[0], r[1]), smp);
            float *q = out + 6 * i;
            q[0] = tc.u; q[1] = tc.v; q[2] = tc.dudx; q[3] = tc.dudy; q[4] = tc.dvdx; q[5] = tc.dvdy;
        }
    }
}
// probes of the environment map: what 0 = evalEnvironment(d) (in 3n -> out 3n), 1 = evalEnvironment with differential directions
// (in 9n: d, rxD, ryD -> out 3n), 2 = pdfEmitterDirect for direction d (in 3n -> out n)
__global__ void k_envmap_probe(DScene sc, int what, uint64_t n, const float *in, float *out) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const DEnvMap &e = *sc.envmap;
        if (what == 2) {
            const float *r = in + 3 * i;
            const DEmitter &em = sc.emitters[sc.envEmitter];
            out[i] = envPdfDirection(e, envToLocal(e, V3(r[0], r[1], r[2]))) * (em.samplingWeight * sc.emitterNormalization);
        } else {
            const float *r = in + (what == 1 ? 9 : 3) * i;
            const Spectrum v = what == 1 ? envEval(e, sc.ewaLut, V3(r[0], r[1], r[2]), true, V3(r[3], r[4], r[5]), V3(r[6], r[7], r[8]))
                                         : envEval(e, sc.ewaLut, V3(r[0], r[1], r[2]), false, V3(0.0f), V3(0.0f));
            out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z;
        }
    }
}
__global__ void k_camera_rays(DScene sc, uint64_t n, const float *pos, float *rays) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        V3 o, d;
        float mint, maxt;
        cameraRay(sc.cam, pos[2 * i], pos[2 * i + 1], o, d, mint, maxt);
        float *r = rays + 8 * i;
        r[0] = o.x; r[1] = o.y; r[2] = o.z; r[3] = mint; r[4] = d.x; r[5] = d.y; r[6] = d.z; r[7] = maxt;
    }
}
__global__ void k_sampler_stream(DScene sc, DRender rp, int px, int py, int sampleIdx, int ndim, float *out) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        PathSampler smp;
        float ax, ay;
        samplerInit(sc, rp, px, py, (uint32_t) sampleIdx, smp, ax, ay);
        if (ndim > 0) out[0] = ax;
        if (ndim > 1) out[1] = ay;
        for (int i = 2; i < ndim; ++i) out[i] = smp.next1D();
    }
}
__global__ void k_splat(DFilter f, int W, int H, uint64_t n, const float *pos, const float *val, float4 *rgba, float *wgt) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const int px = (int) floorf(pos[2 * i]), py = (int) floorf(pos[2 * i + 1]);
        if (px < 0 || py < 0 || px >= W || py >= H) continue;
        filmPut(f, rgba, wgt, W, H, pos[2 * i], pos[2 * i + 1], px, py, V3(val[4 * i], val[4 * i + 1], val[4 * i + 2]), val[4 * i + 3]);
    }
}

// ------------------------------------------------------------------------------------------------
// launchers (host side of this translation unit)
// ------------------------------------------------------------------------------------------------
static size_t traceSmemBytes(const DScene &sc, int block) { // the larger of the two carve-ups of setupTraceMem
    size_t off = (size_t) B2_STACK_DEPTH * block * sizeof(uint32_t);
    off = (off + 127) & ~(size_t) 127;
    off += (size_t) sc.stageNodes * 64 + (size_t) sc.stageTriBytes;
    off = (off + 15) & ~(size_t) 15;
    size_t off8 = 0;
    if (sc.nodes8) {
        off8 = (size_t) B2_STACK8_DEPTH * block * sizeof(uint2);
        off8 = (off8 + 127) & ~(size_t) 127;
        off8 += (size_t) sc.stageNodes8 * 80 + (size_t) sc.stageTriBytes;
        off8 = (off8 + 15) & ~(size_t) 15;
    }
    return (off > off8 ? off : off8) + 16;
}

template <typename K> static int occupancyGrid(K kernel, int block, size_t smem, int numSMs) {
    int perSM = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kernel, block, smem);
    if (perSM < 1) perSM = 1;
    return perSM * numSMs; // a multiple of the SM count: every SM holds the same number of resident CTAs
}

// The opt-in limit is a per-function, process-wide attribute: it is set to one generous bound (never lowered), because
// several committed scenes with different staging sizes share the kernels (a scene committed later must not shrink the
// limit of one committed earlier).  The bound covers the stack (32 KB) + 256 staged nodes + 64 staged triangles.
static void setSmemAttr(const void *fn, size_t smem) {
    const size_t bound = 96 * 1024;
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (smem > bound ? smem : bound));
}

void KernelSet_init(LaunchCfg &cfg, const DScene &sc, int numSMs) {
    cfg.numSMs = numSMs;
    cfg.traceSmem = traceSmemBytes(sc, B2_TRACE_BLOCK);
    setSmemAttr((const void *) k_extend<false>, cfg.traceSmem);
    setSmemAttr((const void *) k_extend<true>, cfg.traceSmem);
    setSmemAttr((const void *) k_occluded, cfg.traceSmem);
    setSmemAttr((const void *) k_trace_rays<false, false>, cfg.traceSmem);
    setSmemAttr((const void *) k_trace_rays<true, false>, cfg.traceSmem);
    setSmemAttr((const void *) k_trace_rays<false, true>, cfg.traceSmem);
    setSmemAttr((const void *) k_trace_rays<true, true>, cfg.traceSmem);
    cfg.flatSmem = (((size_t) sc.stageTriBytes + 15) & ~(size_t) 15) + 16;
    if (sc.rootCount) {
        cfg.gridExtendFlat = occupancyGrid(k_extend_flat<false>, B2_TRACE_BLOCK, cfg.flatSmem, numSMs);
        cfg.gridExtendFlatSort = occupancyGrid(k_extend_flat<true>, B2_TRACE_BLOCK, cfg.flatSmem, numSMs);
        cfg.gridOccludedFlat = occupancyGrid(k_occluded_flat, B2_TRACE_BLOCK, cfg.flatSmem, numSMs);
    }
    cfg.gridExtend = occupancyGrid(k_extend<false>, B2_TRACE_BLOCK, cfg.traceSmem, numSMs);
    cfg.gridExtendSort = occupancyGrid(k_extend<true>, B2_TRACE_BLOCK, cfg.traceSmem, numSMs);
    cfg.gridOccluded = occupancyGrid(k_occluded, B2_TRACE_BLOCK, cfg.traceSmem, numSMs);
    cfg.gridTrace = occupancyGrid(k_trace_rays<false, false>, B2_TRACE_BLOCK, cfg.traceSmem, numSMs);
    setSmemAttr((const void *) k_volstep, cfg.traceSmem);
    cfg.gridVolstep = occupancyGrid(k_volstep, B2_TRACE_BLOCK, cfg.traceSmem, numSMs);
    setSmemAttr((const void *) k_volstep_lockstep, cfg.traceSmem);
    cfg.gridVolLockstep = occupancyGrid(k_volstep_lockstep, B2_TRACE_BLOCK, cfg.traceSmem, numSMs);
    cfg.volLockstep = getenv("B2_VOL_LOCKSTEP") ? atoi(getenv("B2_VOL_LOCKSTEP")) : 1; // measured: 127 vs 95 Msamples/s (smoke 128^3, 512^2 @ 256 spp)
    cfg.gridGenerate = occupancyGrid(k_generate, B2_GEN_BLOCK, 0, numSMs);
    cfg.gridShade[0] = occupancyGrid(k_shade<0>, B2_SHADE_BLOCK, 0, numSMs);
    cfg.gridShade[1] = occupancyGrid(k_shade<1>, B2_SHADE_BLOCK, 0, numSMs);
    cfg.gridShade[2] = occupancyGrid(k_shade<2>, B2_SHADE_BLOCK, 0, numSMs);
    cfg.gridShade[3] = occupancyGrid(k_shade<3>, B2_SHADE_BLOCK, 0, numSMs);
    cfg.gridShade[4] = occupancyGrid(k_shade<-1>, B2_SHADE_BLOCK, 0, numSMs);
    cfg.gridShadeTex = occupancyGrid(k_shade<-1, true>, B2_SHADE_BLOCK, 0, numSMs);
    cfg.gridShadeTexCls[0] = occupancyGrid(k_shade<0, true>, B2_SHADE_BLOCK, 0, numSMs);
    cfg.gridShadeTexCls[1] = occupancyGrid(k_shade<1, true>, B2_SHADE_BLOCK, 0, numSMs);
    cfg.gridShadeTexCls[2] = occupancyGrid(k_shade<2, true>, B2_SHADE_BLOCK, 0, numSMs);
    cfg.gridShadeTexCls[3] = occupancyGrid(k_shade<3, true>, B2_SHADE_BLOCK, 0, numSMs);
}

void launch_generate(const LaunchCfg &cfg, const DScene &sc, const DPool &pool, const DRender &rp, const DFilter &f, cudaStream_t st) {
    k_generate<<<cfg.gridGenerate, B2_GEN_BLOCK, 0, st>>>(sc, pool, rp, f);
    k_publish<<<1, 32, 0, st>>>(pool, rp);
}
void launch_extend(const LaunchCfg &cfg, const DScene &sc, const DPool &pool, const DRender &rp, bool sort, cudaStream_t st) {
    if (sc.rootCount) {
        if (sort) k_extend_flat<true><<<cfg.gridExtendFlatSort, B2_TRACE_BLOCK, cfg.flatSmem, st>>>(sc, pool, rp);
        else k_extend_flat<false><<<cfg.gridExtendFlat, B2_TRACE_BLOCK, cfg.flatSmem, st>>>(sc, pool, rp);
        return;
    }
    if (sort) k_extend<true><<<cfg.gridExtendSort, B2_TRACE_BLOCK, cfg.traceSmem, st>>>(sc, pool, rp);
    else k_extend<false><<<cfg.gridExtend, B2_TRACE_BLOCK, cfg.traceSmem, st>>>(sc, pool, rp);
}
void launch_shade(const LaunchCfg &cfg, const DScene &sc, const DPool &pool, const DRender &rp, int cls, bool queued, cudaStream_t st) {
    // cls 0..3: a specialised instance; cls 4 (queued) or -1 (unqueued): the generic instance
    const uint32_t *q = queued ? pool.matQueue + (size_t) cls * pool.capacity : nullptr;
    const unsigned long long *qc = queued ? pool.counters + (cls < 4 ? CTR_CLASS0 + cls : CTR_CLASSG) : nullptr;
    // scenes with bitmap textures or an environment map: the instances that carry the look-up code (one per BSDF class, so that the
    // class-sorted queues keep a warp on one BSDF: the unsorted generic instance ran 4 of 32 lanes per instruction on mixed materials)
    if (sc.nTextures || sc.envmap) {
        switch (cls) {
            case 0: k_shade<0, true><<<cfg.gridShadeTexCls[0], B2_SHADE_BLOCK, 0, st>>>(sc, pool, rp, q, qc); break;
            case 1: k_shade<1, true><<<cfg.gridShadeTexCls[1], B2_SHADE_BLOCK, 0, st>>>(sc, pool, rp, q, qc); break;
            case 2: k_shade<2, true><<<cfg.gridShadeTexCls[2], B2_SHADE_BLOCK, 0, st>>>(sc, pool, rp, q, qc); break;
            case 3: k_shade<3, true><<<cfg.gridShadeTexCls[3], B2_SHADE_BLOCK, 0, st>>>(sc, pool, rp, q, qc); break;
            default: k_shade<-1, true><<<cfg.gridShadeTex, B2_SHADE_BLOCK, 0, st>>>(sc, pool, rp, q, qc); break;
        }
        return;
    }
    switch (cls) {
        case 0: k_shade<0><<<cfg.gridShade[0], B2_SHADE_BLOCK, 0, st>>>(sc, pool, rp, q, qc); break;
        case 1: k_shade<1><<<cfg.gridShade[1], B2_SHADE_BLOCK, 0, st>>>(sc, pool, rp, q, qc); break;
        case 2: k_shade<2><<<cfg.gridShade[2], B2_SHADE_BLOCK, 0, st>>>(sc, pool, rp, q, qc); break;
        case 3: k_shade<3><<<cfg.gridShade[3], B2_SHADE_BLOCK, 0, st>>>(sc, pool, rp, q, qc); break;
        default:
            k_shade<-1><<<cfg.gridShade[4], B2_SHADE_BLOCK, 0, st>>>(sc, pool, rp, q, qc);
            break;
    }
}
void launch_occluded(const LaunchCfg &cfg, const DScene &sc, const DPool &pool, const DRender &rp, cudaStream_t st) {
    if (sc.rootCount) k_occluded_flat<<<cfg.gridOccludedFlat, B2_TRACE_BLOCK, cfg.flatSmem, st>>>(sc, pool, rp);
    else k_occluded<<<cfg.gridOccluded, B2_TRACE_BLOCK, cfg.traceSmem, st>>>(sc, pool, rp);
}
void launch_volstep(const LaunchCfg &cfg, const DScene &sc, const DPool &pool, const DRender &rp, cudaStream_t st) {
    if (cfg.volLockstep) k_volstep_lockstep<<<cfg.gridVolLockstep, B2_TRACE_BLOCK, cfg.traceSmem, st>>>(sc, pool, rp);
    else k_volstep<<<cfg.gridVolstep, B2_TRACE_BLOCK, cfg.traceSmem, st>>>(sc, pool, rp);
}
void launch_medium_probe(const LaunchCfg &cfg, const DScene &sc, int medium, int what, uint64_t n, const float *in, uint64_t seed, float *out,
                         cudaStream_t st) {
    k_medium_probe<<<cfg.numSMs * 4, 128, 0, st>>>(sc, medium, what, n, in, seed, out);
}
void launch_film_pack(const LaunchCfg &cfg, const float4 *rgba, const float *w, float *out, size_t n, cudaStream_t st) {
    k_film_pack<<<cfg.numSMs * 4, 256, 0, st>>>(rgba, w, out, n);
}
void launch_trace(const LaunchCfg &cfg, const DScene &sc, const float4 *rays, float4 *out, uint64_t n, bool shadow, bool count,
                  unsigned long long *counters, const uint32_t *order, cudaStream_t st) {
    const int g = cfg.gridTrace;
    if (shadow) {
        if (count) k_trace_rays<true, true><<<g, B2_TRACE_BLOCK, cfg.traceSmem, st>>>(sc, rays, out, n, counters, order);
        else k_trace_rays<true, false><<<g, B2_TRACE_BLOCK, cfg.traceSmem, st>>>(sc, rays, out, n, counters, order);
    } else {
        if (count) k_trace_rays<false, true><<<g, B2_TRACE_BLOCK, cfg.traceSmem, st>>>(sc, rays, out, n, counters, order);
        else k_trace_rays<false, false><<<g, B2_TRACE_BLOCK, cfg.traceSmem, st>>>(sc, rays, out, n, counters, order);
    }
}
// counting sort of ray tickets by (entry cell, direction cell): keys n, hist B2_NBINS + 1 (left holding the bin END offsets), order n
void launch_bin(const LaunchCfg &cfg, const DScene &sc, const DPool &pool, const float4 *rays, uint32_t n, uint32_t *keys, uint32_t *hist,
                uint32_t *order, cudaStream_t st) {
    cudaMemsetAsync(hist, 0, (B2_NBINS + 1) * sizeof(uint32_t), st);
    k_bin_count<<<cfg.numSMs * 8, 256, 0, st>>>(sc, pool, rays, n, keys, hist);
    k_bin_scan<<<1, 1024, 0, st>>>(hist);
    k_bin_scatter<<<cfg.numSMs * 8, 256, 0, st>>>(n, keys, hist, order);
}
void launch_bsdf_eval(const LaunchCfg &cfg, const DScene &sc, int mat, uint64_t n, const float *wi, const float *wo, float *rgb, float *pdf, cudaStream_t st) {
    k_bsdf_eval<<<cfg.numSMs * 2, 128, 0, st>>>(sc, mat, n, wi, wo, rgb, pdf);
}
void launch_bsdf_sample(const LaunchCfg &cfg, const DScene &sc, int mat, uint64_t n, const float *wi, const float *samples, float *out, cudaStream_t st) {
    k_bsdf_sample<<<cfg.numSMs * 2, 128, 0, st>>>(sc, mat, n, wi, samples, out);
}
void launch_emitter_direct(const LaunchCfg &cfg, const DScene &sc, uint64_t n, const float *ref, const float *samples, float *out, cudaStream_t st) {
    k_emitter_direct<<<cfg.numSMs * 2, 128, 0, st>>>(sc, n, ref, samples, out);
}
void launch_texture_probe(const LaunchCfg &cfg, const DScene &sc, int what, int tex, int hasPartials, float diffScale, uint64_t n, const float *in, float *out,
                          cudaStream_t st) {
    k_texture_probe<<<cfg.numSMs * 4, 128, 0, st>>>(sc, what, tex, hasPartials, diffScale, n, in, out);
}
void launch_envmap_probe(const LaunchCfg &cfg, const DScene &sc, int what, uint64_t n, const float *in, float *out, cudaStream_t st) {
    k_envmap_probe<<<cfg.numSMs * 4, 128, 0, st>>>(sc, what, n, in, out);
}
void launch_camera_rays(const LaunchCfg &cfg, const DScene &sc, uint64_t n, const float *pos, float *rays, cudaStream_t st) {
    k_camera_rays<<<cfg.numSMs * 2, 128, 0, st>>>(sc, n, pos, rays);
}
void launch_sampler_stream(const DScene &sc, const DRender &rp, int px, int py, int sampleIdx, int ndim, float *out, cudaStream_t st) {
    k_sampler_stream<<<1, 32, 0, st>>>(sc, rp, px, py, sampleIdx, ndim, out);
}
void launch_splat(const LaunchCfg &cfg, const DFilter &f, int W, int H, uint64_t n, const float *pos, const float *val, float4 *rgba, float *wgt, cudaStream_t st) {
    k_splat<<<cfg.numSMs * 2, 128, 0, st>>>(f, W, H, n, pos, val, rgba, wgt);
}

} // namespace B2_KNS
} // namespace b2
