// TEST INFRASTRUCTURE — NOT PRODUCT CODE.  What the CPU functional emulator cannot offer: CUDA IPC is refused, so
// multi-rank solves take the ncclSend/ncclRecv path (run the ranks with FI_B200_P2P=0, which skips the attempt).
#include <cuda_runtime.h>

extern "C" {
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return cudaErrorNotSupported; }
cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
}
