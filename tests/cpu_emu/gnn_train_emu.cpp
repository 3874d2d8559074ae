// csrc/gnn_train.cuh compiled for the host (see cuda_emu.h) behind the signature of the C ABI entry points
// deepaco_gnn_train_forward / deepaco_gnn_train_backward.  Test infrastructure only.
#include "cuda_emu.h"

#include "../../deepaco_b200/csrc/gnn_train.cuh"
#include "../../deepaco_b200/csrc/gnn_train_args.h"

using namespace deepaco::gnnt;

extern "C" const char* emu_gnn_train_forward(const deepaco_gnn_train_args* a, int threads) {
    TrainParams p;
    if (const char* err = gnn_train_params(a, kTrainForward, p)) return err;
    memset(a->sync_ws, 0, sizeof(uint32_t) * a->n_instances);
    emu::launch(gnn_group_forward_kernel<true>, p, a->n_instances, a->ctas_per_instance, threads, smem_floats_fwd(threads) * 4);
    return nullptr;
}

extern "C" const char* emu_gnn_train_backward(const deepaco_gnn_train_args* a, int threads) {
    TrainParams p;
    if (const char* err = gnn_train_params(a, kTrainBackward, p)) return err;
    memset(a->sync_ws, 0, sizeof(uint32_t) * a->n_instances);
    emu::launch(gnn_train_backward_kernel, p, a->n_instances, a->ctas_per_instance, threads, smem_floats_bwd(threads) * 4);
    return nullptr;
}

extern "C" const char* emu_gnn_forward_group(const deepaco_gnn_train_args* a, int threads) {
    TrainParams p;
    if (const char* err = gnn_train_params(a, kEvalForward, p)) return err;
    memset(a->sync_ws, 0, sizeof(uint32_t) * a->n_instances);
    emu::launch(gnn_group_forward_kernel<false>, p, a->n_instances, a->ctas_per_instance, threads, smem_floats_fwd(threads) * 4);
    return nullptr;
}
