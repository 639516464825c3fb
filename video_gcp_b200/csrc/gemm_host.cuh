// Host-side helpers for the GEMM kernels: TMA tensor-map construction (driver entry point fetched at
// run time, so the library has no link-time dependency on libcuda) and launchers.
#pragma once
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include "gemm.cuh"

namespace gcp {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// bf16 row-major [rows][ld] array viewed as a 2-D tensor (cols x rows); box = 64 columns x box_rows rows,
// 128B swizzle (64 bf16 = 128 B per box row), out-of-bounds reads return zeros.
inline int make_tmap_bf16(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                          uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (enc == nullptr) {
        gcp_set_error("cuTensorMapEncodeTiled entry point not available");
        return -1;
    }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        gcp_set_error("cuTensorMapEncodeTiled failed: CUresult %d (rows %llu cols %llu ld %llu box_rows %u)", (int)r,
                      (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows);
        return -1;
    }
    return 0;
}

// Cluster size used for `rows`: the largest of {max_cluster, ..., 2, 1} that divides the number of M tiles.
inline int gemm_cluster_size(int rows, int max_cluster) {
    const int tiles_m = rows / GEMM_BM;
    int c = max_cluster;
    while (c > 1 && (tiles_m % c)) c >>= 1;
    return c < 1 ? 1 : c;
}

// Programmatic dependent launch of the GEMM chain (the kernel's prologue overlaps its predecessor's tail; see
// pdl_wait() in common.cuh).  GCPB200_NO_PDL=1 turns it off for A/B measurements; gemm_pdl_suspend() is used around
// stream capture.
inline int& gemm_pdl_suspended() {
    static int s = 0;
    return s;
}
inline bool gemm_pdl_enabled() {
#ifdef GCPB200_VERIFY
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("GCPB200_NO_PDL");
        on = (e != nullptr && e[0] == '1') ? 0 : 1;
    }
    return on == 1 && gemm_pdl_suspended() == 0;
#else
    return gemm_pdl_suspended() == 0;
#endif
}

// CTA-pair MMAs for the BN = 256 GEMMs whose launch uses 2-CTA clusters (GCPB200_NO_PAIR=1 in the verification build
// switches back to two independent M = 128 CTAs with a multicast weight tile, for A/B measurements).
inline bool gemm_pair_enabled() {
#ifdef GCPB200_VERIFY
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("GCPB200_NO_PAIR");
        on = (e != nullptr && e[0] == '1') ? 0 : 1;
    }
    return on == 1;
#else
    return true;
#endif
}

template <int BN, int EPI, bool PAIR = false>
int launch_gemm_tc(const GemmArgs& a, cudaStream_t st, int num_sms, int cluster) {
    using Cfg = GemmCfg<BN, PAIR>;
    // function attributes and occupancy are per device: one slot per device ordinal (a process normally drives one GPU,
    // but nothing here should break when it owns several contexts)
    static bool configured_dev[64] = {false};
    static int max_clusters_dev[64][9] = {{0}};   // co-resident clusters per cluster size
    int dev = 0;
    GCP_CUDA_CHECK(cudaGetDevice(&dev));
    dev &= 63;
    int* max_clusters = max_clusters_dev[dev];
    if (!configured_dev[dev]) {
        GCP_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            Cfg::SMEM_BYTES));
        configured_dev[dev] = true;
    }
    if (a.rows % GEMM_BM || a.N % BN || a.K % GEMM_BK || (a.rows / GEMM_BM) % cluster) {
        gcp_set_error("gemm: bad shape rows %d N %d K %d (BN %d, cluster %d)", a.rows, a.N, a.K, BN, cluster);
        return -1;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = gemm_pdl_enabled() ? 2 : 1;
    if (max_clusters[cluster] == 0) {
        cfg.gridDim = dim3(num_sms / cluster * cluster);
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, gemm_tc_kernel<BN, EPI, PAIR>, &cfg) != cudaSuccess || n <= 0) n = num_sms / cluster;
        max_clusters[cluster] = n;
    }
    const int n_work = (a.rows / GEMM_BM / cluster) * (a.N / BN);
    const int n_clusters = n_work < max_clusters[cluster] ? n_work : max_clusters[cluster];
    cfg.gridDim = dim3(n_clusters * cluster);
    GCP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, EPI, PAIR>, a));
    return 0;
}

#ifdef GCPB200_VERIFY
template <int EPI>
int launch_gemm_ref(const GemmArgs& a, int BN, cudaStream_t st) {
    dim3 grid(a.rows / GEMM_BM, a.N / BN);
    gemm_ref_kernel<EPI><<<grid, 128, 0, st>>>(a, BN);
    GCP_CUDA_CHECK(cudaGetLastError());
    return 0;
}
#endif

// dispatch on (BN, EPI); use_ref selects the SIMT verification kernels, which only the GCPB200_VERIFY build contains
inline int launch_gemm(const GemmArgs& a, int BN, int epi, bool use_ref, cudaStream_t st, int num_sms, int cluster = 1) {
#ifdef GCPB200_VERIFY
    if (use_ref) {
        switch (epi) {
            case EPI_LINEAR: return launch_gemm_ref<EPI_LINEAR>(a, BN, st);
            case EPI_GN: return launch_gemm_ref<EPI_GN>(a, BN, st);
            case EPI_REPARAM: return launch_gemm_ref<EPI_REPARAM>(a, BN, st);
            case EPI_LSTM: return launch_gemm_ref<EPI_LSTM>(a, BN, st);
        }
    }
#endif
    (void)use_ref;
    if (BN == 128) {
        switch (epi) {
            case EPI_LINEAR: return launch_gemm_tc<128, EPI_LINEAR>(a, st, num_sms, cluster);
            case EPI_GN: return launch_gemm_tc<128, EPI_GN>(a, st, num_sms, cluster);
        }
    } else if (BN == 256) {
        if (cluster == 2 && gemm_pair_enabled()) {
            // an even number of 128-row tiles: CTA pairs run M = 256 MMAs (cta_group::2)
            switch (epi) {
                case EPI_LINEAR: return launch_gemm_tc<256, EPI_LINEAR, true>(a, st, num_sms, cluster);
                case EPI_REPARAM: return launch_gemm_tc<256, EPI_REPARAM, true>(a, st, num_sms, cluster);
                case EPI_LSTM: return launch_gemm_tc<256, EPI_LSTM, true>(a, st, num_sms, cluster);
            }
        }
        switch (epi) {
            case EPI_LINEAR: return launch_gemm_tc<256, EPI_LINEAR>(a, st, num_sms, cluster);
            case EPI_REPARAM: return launch_gemm_tc<256, EPI_REPARAM>(a, st, num_sms, cluster);
            case EPI_LSTM: return launch_gemm_tc<256, EPI_LSTM>(a, st, num_sms, cluster);
        }
    }
    gcp_set_error("gemm: unsupported BN %d / epilogue %d", BN, epi);
    return -1;
}

}  // namespace gcp
