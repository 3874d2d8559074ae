// C ABI of the training-mode heuristic network (kernels in gnn_train.cuh): one group of CTAs per graph.
#include "gnn_train.cuh"
#include "gnn_train_args.h"
#include "host_util.h"

using namespace deepaco;
using namespace deepaco::gnnt;

// groups of <= 8 CTAs: one thread-block cluster per graph.  Larger groups: cooperative launches (co-residency is what
// makes the arrival-counter barrier safe), as many graphs per launch as fit on the device at once.
template <class Kernel>
static int launch_groups(Kernel kernel, TrainParams p, int n_instances, int ctas, int threads, size_t smem, cudaStream_t st) {
    DACO_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (p.grid_ctas == 0) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(n_instances * ctas));
        cfg.blockDim = dim3((unsigned)threads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)ctas;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        DACO_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, p));
        DACO_CHECK_LAUNCH();
        return DEEPACO_OK;
    }
    const DeviceInfo* di = device_info();
    DACO_CHECK_ARG(di != nullptr, "deepaco_gnn_train: no CUDA device");
    int per_sm = 0;
    DACO_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    const int resident = per_sm * di->sm_count;
    DACO_CHECK_ARG(resident >= ctas, "deepaco_gnn_train: %d CTAs per graph cannot be co-resident on this device (%d fit)", ctas, resident);
    const int per_launch = resident / ctas;
    DACO_CHECK_CUDA(cudaMemsetAsync(p.sync_ctr, 0, sizeof(unsigned) * (size_t)n_instances, st));
    for (int b0 = 0; b0 < n_instances; b0 += per_launch) {
        const int nb = n_instances - b0 < per_launch ? n_instances - b0 : per_launch;
        p.b0 = b0;
        void* args[] = {&p};
        DACO_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)kernel, dim3((unsigned)(nb * ctas)), dim3((unsigned)threads), args, smem, st));
        DACO_CHECK_LAUNCH();
    }
    return DEEPACO_OK;
}

extern "C" int deepaco_gnn_train_forward(const deepaco_gnn_train_args* a, void* stream) {
    TrainParams p;
    if (const char* err = gnn_train_params(a, kTrainForward, p)) DACO_CHECK_ARG(false, "deepaco_gnn_train_forward: %s", err);
    const int threads = a->n_edges >= 2048 * a->ctas_per_instance ? 512 : 256;
    return launch_groups(gnn_group_forward_kernel<true>, p, a->n_instances, a->ctas_per_instance, threads,
                            smem_floats_fwd(threads) * 4, (cudaStream_t)stream);
}

extern "C" int deepaco_gnn_train_backward(const deepaco_gnn_train_args* a, void* stream) {
    TrainParams p;
    if (const char* err = gnn_train_params(a, kTrainBackward, p)) DACO_CHECK_ARG(false, "deepaco_gnn_train_backward: %s", err);
    const int threads = 256;
    return launch_groups(gnn_train_backward_kernel, p, a->n_instances, a->ctas_per_instance, threads,
                            smem_floats_bwd(threads) * 4, (cudaStream_t)stream);
}

// eval-mode forward (running-statistics BatchNorm) by a group of CTAs per graph: the low-latency form of
// deepaco_gnn_forward for one or a few instances (Net.forward in eval mode, tsp/net.py:84-88)
extern "C" int deepaco_gnn_forward_group(const deepaco_gnn_train_args* a, void* stream) {
    TrainParams p;
    if (const char* err = gnn_train_params(a, kEvalForward, p)) DACO_CHECK_ARG(false, "deepaco_gnn_forward_group: %s", err);
    const int threads = a->n_edges >= 2048 * a->ctas_per_instance ? 512 : 256;
    return launch_groups(gnn_group_forward_kernel<false>, p, a->n_instances, a->ctas_per_instance, threads,
                         smem_floats_fwd(threads) * 4, (cudaStream_t)stream);
}
