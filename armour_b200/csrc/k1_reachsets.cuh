// K1/K2 reach-set construction — placeholder until the kernel lands (next commit).
#pragma once
#include <cuda_runtime.h>

#include "../../include/armour_b200.h"
#include "layout.h"

namespace armour {

struct K1Scratch {
    void* arena = nullptr;
};
inline cudaError_t k1_scratch_create(K1Scratch*, const armour_config&, const RobotConstants&, cudaStream_t) {
    return cudaSuccess;
}
inline void k1_scratch_destroy(K1Scratch*) {}
inline cudaError_t launch_reachsets(const Batch&, K1Scratch&, cudaStream_t, int* nlaunch) {
    *nlaunch = 0;
    return cudaErrorNotSupported;
}

}  // namespace armour
