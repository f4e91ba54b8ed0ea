// Micro-benchmarks / layout checks for the cluster recurrence kernel (round 2): not product code, not a test.
//   G  distributed-shared-memory exchange with cp.async.bulk (shared::cta -> shared::cluster, complete_tx on the peer's mbarrier)
//   H  UMMA shared-memory descriptor without swizzle (K-major "interleave": 8-row x 16-byte core matrices), checked numerically
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../audio_only_speech_separation_b200/csrc ubench_cluster.cu -o ubench_cluster
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc5_common.cuh"

using namespace dp;

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);    \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (!done) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (!done && ++spins > SPIN_LIMIT) __trap();
    }
}
__device__ __forceinline__ void bulk_copy_dsmem(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst_cluster), "r"(src_cta),
                 "r"(bytes), "r"(bar_cluster)
                 : "memory");
}

// ------------------------------------------------------------------------------------------------ G
template <int CS>
__global__ void __launch_bounds__(256, 1) dsmem_bulk_kernel(int bytes, int ncopies, int steps, int write_src, long long* out, int* check) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar[2];
    const int tid = threadIdx.x;
    const uint32_t rank = cluster_rank();
    uint8_t* src = smem;                    // [2][bytes]
    uint8_t* dst = smem + 2 * 16384;        // [2][CS][bytes]  (bytes <= 16384)
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    cluster_sync_all();
    const int nvec = bytes / 16;
    const int part = bytes / ncopies;
    int bad = 0;
    long long t0 = clock64();
    for (int s = 0; s < steps; ++s) {
        const int b = s & 1;
        uint8_t* sb = src + b * 16384;
        if (write_src) {
            for (int i = tid; i < nvec; i += 256) *reinterpret_cast<uint4*>(sb + i * 16) = make_uint4(s, i, rank, 7);
            proxy_fence_async();
        }
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&bar[b], (uint32_t)((CS - 1) * bytes));
            for (int pr = 1; pr < CS; ++pr) {
                const uint32_t peer = (rank + pr) % CS;
                const uint32_t d = mapa(smem_u32(dst + (b * CS + rank) * 16384), peer);
                const uint32_t mb = mapa(smem_u32(&bar[b]), peer);
                for (int c = 0; c < ncopies; ++c) bulk_copy_dsmem(d + c * part, smem_u32(sb) + c * part, (uint32_t)part, mb);
            }
        }
        mbar_wait_cluster(&bar[b], (s >> 1) & 1);
        if (write_src && s == steps - 1) {   // check the received data of the last step
            for (int pr = 1; pr < CS; ++pr) {
                const uint32_t peer = (rank + pr) % CS;
                for (int i = tid; i < nvec; i += 256) {
                    const uint4 v = *reinterpret_cast<const uint4*>(dst + (b * CS + peer) * 16384 + i * 16);
                    if (v.x != (uint32_t)s || v.y != (uint32_t)i || v.z != peer || v.w != 7) bad = 1;
                }
            }
        }
    }
    long long t1 = clock64();
    if (tid == 0) out[blockIdx.x] = t1 - t0;
    if (bad) atomicAdd(check, 1);
    cluster_sync_all();
}

// ------------------------------------------------------------------------------------------------ H
// no-swizzle K-major descriptor: core matrix = 8 rows x 16 bytes (contiguous 128 B); SBO = distance between 8-row groups,
// LBO = distance between the two 16-byte K chunks of one K = 16 instruction
__device__ __forceinline__ uint64_t desc_nosw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// A[r][kk] = r + 1 (others 0) in the interleaved layout [chunk][row][16 B] (LBO = rows*16, SBO = 128), B[n][kk] = n + 1 (SW128),
// two K = 16 instructions (K = 32) -> D[r][n] = (r + 1)(n + 1)
__global__ void __launch_bounds__(128, 1) nosw_kernel(int kk, int a_is_b, float* dump) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    __syncthreads();
    uint8_t* il = smem;            // interleaved operand: [4 chunks][128 rows][16 B] = 8 KB
    uint8_t* sw = smem + 16384;    // SW128 operand: [128 rows][128 B]
    {
        const int r = tid;
        *reinterpret_cast<__nv_bfloat16*>(il + (kk >> 3) * 2048 + r * 16 + (kk & 7) * 2) = __float2bfloat16_rn((float)(r + 1));
        *reinterpret_cast<__nv_bfloat16*>(sw + r * 128 + (((kk >> 3) ^ (r & 7)) << 4) + (kk & 7) * 2) = __float2bfloat16_rn((float)(r + 1));
    }
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
    if (warp == 0) tmem_alloc(&slot, 512);
    proxy_fence_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 1) {
        const uint32_t idesc = idesc_bf16(128, 128, 0, 0);
        for (int k = 0; k < 2; ++k) {
            const uint64_t d_il = desc_nosw(smem_u32(il) + k * 2 * 2048, 2048, 128);
            const uint64_t d_sw = desc_sw128(smem_u32(sw), 16, 1024) + (uint64_t)(k * 2);
            if (a_is_b) umma_w(tmem, d_sw, d_il, idesc, k != 0);   // interleaved operand as B
            else umma_w(tmem, d_il, d_sw, idesc, k != 0);          // interleaved operand as A
        }
        umma_commit_w(&bar);
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c = 0; c < 128; c += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
        for (int i = 0; i < 32; ++i) dump[(warp * 32 + lane) * 128 + c + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ E2: tight issue loop, A operand in tensor memory
__device__ __forceinline__ void umma_ts_x8(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc) {
    asm volatile(
        "{\n .reg .pred q;\n .reg .b64 b;\n .reg .b32 a;\n elect.sync _|q, 0xffffffff;\n mov.b32 a, %1;\n mov.b64 b, %2;\n"
        " @q tcgen05.mma.cta_group::1.kind::f16 [%0], [a], b, %3, 1;\n add.u32 a, a, 8;\n add.u64 b, b, 2;\n"
        " @q tcgen05.mma.cta_group::1.kind::f16 [%0], [a], b, %3, 1;\n add.u32 a, a, 8;\n add.u64 b, b, 2;\n"
        " @q tcgen05.mma.cta_group::1.kind::f16 [%0], [a], b, %3, 1;\n add.u32 a, a, 8;\n add.u64 b, b, 2;\n"
        " @q tcgen05.mma.cta_group::1.kind::f16 [%0], [a], b, %3, 1;\n add.u32 a, a, 8;\n sub.u64 b, b, 6;\n"
        " @q tcgen05.mma.cta_group::1.kind::f16 [%0], [a], b, %3, 1;\n add.u32 a, a, 8;\n add.u64 b, b, 2;\n"
        " @q tcgen05.mma.cta_group::1.kind::f16 [%0], [a], b, %3, 1;\n add.u32 a, a, 8;\n add.u64 b, b, 2;\n"
        " @q tcgen05.mma.cta_group::1.kind::f16 [%0], [a], b, %3, 1;\n add.u32 a, a, 8;\n add.u64 b, b, 2;\n"
        " @q tcgen05.mma.cta_group::1.kind::f16 [%0], [a], b, %3, 1;\n}\n" ::"r"(d), "r"(a), "l"(db), "r"(idesc)
        : "memory");
}
__device__ __forceinline__ void umma_ss_x8(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc) {
    asm volatile(
        "{\n .reg .pred q;\n .reg .b64 a, b;\n elect.sync _|q, 0xffffffff;\n mov.b64 a, %1;\n mov.b64 b, %2;\n"
        " @q tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, 1;\n add.u64 a, a, 2;\n add.u64 b, b, 2;\n"
        " @q tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, 1;\n add.u64 a, a, 2;\n add.u64 b, b, 2;\n"
        " @q tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, 1;\n add.u64 a, a, 2;\n add.u64 b, b, 2;\n"
        " @q tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, 1;\n sub.u64 a, a, 6;\n sub.u64 b, b, 6;\n"
        " @q tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, 1;\n add.u64 a, a, 2;\n add.u64 b, b, 2;\n"
        " @q tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, 1;\n add.u64 a, a, 2;\n add.u64 b, b, 2;\n"
        " @q tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, 1;\n add.u64 a, a, 2;\n add.u64 b, b, 2;\n"
        " @q tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, 1;\n}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc)
        : "memory");
}
// mode 0: all SS ; 1: all TS ; 2: TS and SS alternating blocks of 8 ; nacc accumulators used round-robin per block of 8 (column offset 256 + i*N)
__global__ void __launch_bounds__(128, 1) mma_tight2_kernel(int M, int N, int rounds, int mode, int nacc, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
    if (warp == 0) tmem_alloc(&slot, 512);
    proxy_fence_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    {   // zero the A region of tensor memory (columns 0..255)
        uint32_t v[32];
        for (int i = 0; i < 32; ++i) v[i] = 0;
        for (int c = 0; c < 256; c += 32) tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) {
        const uint32_t idesc = idesc_bf16(M, N, 0, 0);
        const uint64_t da = desc_sw128(smem_u32(smem), 16, 1024);
        const uint64_t db = desc_sw128(smem_u32(smem + 16384), 16, 1024);
        long long t0 = clock64();
        for (int r = 0; r < rounds; ++r) {
            const uint32_t d = tmem + 256 + (uint32_t)((r % nacc) * N);
            const bool ts = mode == 1 || (mode == 2 && (r & 1));
            if (ts) umma_ts_x8(d, tmem + (uint32_t)((r & 3) * 64), db, idesc);
            else umma_ss_x8(d, da, db, idesc);
        }
        umma_commit_w(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if ((tid & 31) == 0) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ I: fragment layout of tcgen05.ld.16x256b.x4
// D[r][n] = (r + 1)(n + 1) as in H; every warp dumps its 16 registers of the loads at lane offsets 0 and 16 of its quadrant
__global__ void __launch_bounds__(128, 1) frag_kernel(float* dump) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    __syncthreads();
    uint8_t* il = smem;
    uint8_t* sw = smem + 16384;
    {
        const int r = tid;
        *reinterpret_cast<__nv_bfloat16*>(il + r * 16) = __float2bfloat16_rn((float)(r + 1));
        *reinterpret_cast<__nv_bfloat16*>(sw + r * 128 + ((0 ^ (r & 7)) << 4)) = __float2bfloat16_rn((float)(r + 1));
    }
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
    if (warp == 0) tmem_alloc(&slot, 512);
    proxy_fence_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 1) {
        umma_w(tmem, desc_nosw(smem_u32(il), 2048, 128), desc_sw128(smem_u32(sw), 16, 1024), idesc_bf16(128, 32, 0, 0), 0);
        umma_commit_w(&bar);
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int half = 0; half < 2; ++half) {
        uint32_t r[16];
        const uint32_t addr = tmem + ((uint32_t)(warp * 32 + half * 16) << 16);
        asm volatile(
            "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
              "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(addr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        for (int i = 0; i < 16; ++i) dump[((warp * 2 + half) * 32 + lane) * 16 + i] = __uint_as_float(r[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
    long long* out;
    int* check;
    CK(cudaMalloc(&out, 1024 * sizeof(long long)));
    CK(cudaMalloc(&check, sizeof(int)));
    std::vector<long long> h(1024);

    printf("== H: no-swizzle K-major descriptor (interleaved [chunk][row][16B]) as A / as B\n");
    {
        const int SMEM = 16384 + 32768 + 1024;
        CK(cudaFuncSetAttribute(nosw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        float* dump;
        CK(cudaMalloc(&dump, 128 * 128 * sizeof(float)));
        std::vector<float> hd(128 * 128);
        for (int a_is_b = 0; a_is_b < 2; ++a_is_b)
            for (int kk : {0, 5, 9, 17, 30}) {
                nosw_kernel<<<1, 128, SMEM>>>(kk, a_is_b, dump);
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(hd.data(), dump, hd.size() * sizeof(float), cudaMemcpyDeviceToHost));
                int bad = 0;
                for (int r = 0; r < 128; ++r)
                    for (int n = 0; n < 128; ++n)
                        if (hd[r * 128 + n] != (float)((r + 1) * (n + 1))) ++bad;
                printf("H interleaved-as-%s kk=%2d : %s (%d mismatches; D[1][2]=%g D[127][127]=%g)\n", a_is_b ? "B" : "A", kk, bad ? "FAIL" : "ok", bad,
                       hd[1 * 128 + 2], hd[127 * 128 + 127]);
            }
    }

    printf("== I: tcgen05.ld.16x256b.x4 fragment layout (row,col) per register, warp 1 (quadrant lanes 32..63)\n");
    {
        const int SMEM = 16384 + 32768 + 1024;
        CK(cudaFuncSetAttribute(frag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        float* dump;
        CK(cudaMalloc(&dump, 4 * 2 * 32 * 16 * sizeof(float)));
        std::vector<float> hd(4 * 2 * 32 * 16);
        frag_kernel<<<1, 128, SMEM>>>(dump);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hd.data(), dump, hd.size() * sizeof(float), cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int w = 0; w < 4; ++w)
            for (int half = 0; half < 2; ++half)
                for (int l = 0; l < 32; ++l)
                    for (int i = 0; i < 16; ++i) {
                        // expected: row = 32w + 16*half + (l >> 2) + 8*((i >> 1) & 1), col = 8*(i >> 2) + 2*(l & 3) + (i & 1)
                        const int row = 32 * w + 16 * half + (l >> 2) + 8 * ((i >> 1) & 1), col = 8 * (i >> 2) + 2 * (l & 3) + (i & 1);
                        const float v = hd[((w * 2 + half) * 32 + l) * 16 + i];
                        if (v != (float)((row + 1) * (col + 1))) ++bad;
                    }
        printf("I expected-layout mismatches: %d\n", bad);
        for (int l : {0, 1, 5, 31}) {
            printf("I warp 1 half 0 lane %2d:", l);
            for (int i = 0; i < 16; ++i) {
                const float v = hd[((1 * 2 + 0) * 32 + l) * 16 + i];
                int fr = -1, fc = -1;
                for (int r = 0; r < 128 && fr < 0; ++r)
                    for (int c = 0; c < 32; ++c)
                        if (v == (float)((r + 1) * (c + 1)) && r >= 32 && r < 64) { fr = r; fc = c; break; }
                printf(" r%d=(%d,%d)", i, fr, fc);
            }
            printf("\n");
        }
    }
    printf("== E2: tight issue loop (8 MMAs per elect), SS / TS / mixed, cycles per MMA\n");
    {
        const int SMEM = 16384 + 32768 + 1024;
        CK(cudaFuncSetAttribute(mma_tight2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        if (getenv("UB_E2")) for (int mode = 0; mode < 3; ++mode)
            for (int M : {64, 128})
                for (int N : {16, 32, 48, 64})
                    for (int nacc : {1, 4}) {
                        if (N * nacc > 256) continue;
                        mma_tight2_kernel<<<148, 128, SMEM>>>(M, N, 256, mode, nacc, out);
                        CK(cudaDeviceSynchronize());
                        mma_tight2_kernel<<<148, 128, SMEM>>>(M, N, 256, mode, nacc, out);
                        CK(cudaDeviceSynchronize());
                        CK(cudaMemcpy(h.data(), out, 148 * sizeof(long long), cudaMemcpyDeviceToHost));
                        long long mx = 0;
                        for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
                        printf("E2 mode=%s M=%3d N=%3d nacc=%d : %.1f cyc/MMA\n", mode == 0 ? "SS" : mode == 1 ? "TS" : "mix", M, N, nacc, (double)mx / (256.0 * 8));
                    }
    }
    printf("== G: DSMEM exchange by cp.async.bulk + complete_tx, cycles per step\n");
    {
        const int DS = 2 * 16384 + 2 * 4 * 16384 + 1024;
        CK(cudaFuncSetAttribute(dsmem_bulk_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, DS));
        CK(cudaFuncSetAttribute(dsmem_bulk_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, DS));
        CK(cudaFuncSetAttribute(dsmem_bulk_kernel<4>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        if (getenv("UB_G")) for (int cs : {2, 4})
            for (int ws : {0, 1})
                for (int bytes : {16, 2048, 4096, 8192, 16384})
                    for (int nc : {1, 2, 8}) {
                        if (bytes == 16 && nc > 1) continue;
                        cudaLaunchConfig_t cfg = {};
                        cfg.gridDim = dim3(cs == 4 ? 128 : 148);
                        cfg.blockDim = dim3(256);
                        cfg.dynamicSmemBytes = DS;
                        cudaLaunchAttribute at[1];
                        at[0].id = cudaLaunchAttributeClusterDimension;
                        at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                        cfg.attrs = at; cfg.numAttrs = 1;
                        CK(cudaMemset(check, 0, sizeof(int)));
                        for (int rep = 0; rep < 2; ++rep) {
                            if (cs == 4) CK(cudaLaunchKernelEx(&cfg, dsmem_bulk_kernel<4>, bytes, nc, 200, ws, out, check));
                            else CK(cudaLaunchKernelEx(&cfg, dsmem_bulk_kernel<2>, bytes, nc, 200, ws, out, check));
                            CK(cudaDeviceSynchronize());
                        }
                        const int grid = cfg.gridDim.x;
                        int bad = 0;
                        CK(cudaMemcpy(&bad, check, sizeof(int), cudaMemcpyDeviceToHost));
                        CK(cudaMemcpy(h.data(), out, grid * sizeof(long long), cudaMemcpyDeviceToHost));
                        long long mx = 0, mn = 1ll << 60;
                        for (int i = 0; i < grid; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
                        printf("G cluster=%d write_src=%d bytes_per_peer=%5d copies_per_peer=%d : max %.0f min %.0f cyc/step%s\n", cs, ws, bytes, nc,
                               (double)mx / 200, (double)mn / 200, bad ? "  DATA MISMATCH" : "");
                    }
        // how many 4-CTA clusters are co-resident
        int ncl = 0;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(148); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = DS;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, dsmem_bulk_kernel<4>, &cfg);
        printf("G cudaOccupancyMaxActiveClusters(cluster 4, %d B smem): %d (%s)\n", DS, ncl, cudaGetErrorString(e));
        cfg.dynamicSmemBytes = 220 * 1024;
        CK(cudaFuncSetAttribute(dsmem_bulk_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        e = cudaOccupancyMaxActiveClusters(&ncl, dsmem_bulk_kernel<4>, &cfg);
        printf("G cudaOccupancyMaxActiveClusters(cluster 4, 220 KB smem): %d (%s)\n", ncl, cudaGetErrorString(e));
    }
    printf("done\n");
    return 0;
}
