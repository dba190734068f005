// __constant__ block shared by every kernel of the library (single translation unit: kernels.cu).
#pragma once
#ifndef ARMOUR_EMU
#include <cuda_runtime.h>
#endif

#include "robot_constants.h"

namespace armour {

__constant__ RobotConstants c_robot;
__constant__ unsigned char c_combA[NCOMB];  // generator-pair enumeration (0,1),(0,2)...(7,8)
__constant__ unsigned char c_combB[NCOMB];  // reference KPR/CollisionChecking.cu:26-39

#ifndef ARMOUR_EMU
inline cudaError_t upload_constants(const RobotConstants& rc, cudaStream_t stream) {
    unsigned char a[NCOMB], b[NCOMB];
    int ai = 0, bi = 1;
    for (int i = 0; i < NCOMB; i++) {
        a[i] = (unsigned char)ai;
        b[i] = (unsigned char)bi;
        if (bi < 8) {
            bi++;
        } else {
            ai++;
            bi = ai + 1;
        }
    }
    cudaError_t e = cudaMemcpyToSymbolAsync(c_robot, &rc, sizeof(rc), 0, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbolAsync(c_combA, a, sizeof(a), 0, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbolAsync(c_combB, b, sizeof(b), 0, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(stream);  // a, b are stack temporaries
}
#endif

}  // namespace armour
