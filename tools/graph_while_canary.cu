// Canary for the CUDA-graph WHILE node exactly as icp.cu builds it (explicit kernel nodes, device-side condition).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/canary tools/graph_while_canary.cu && /tmp/canary
#include <cuda_runtime.h>
#include <cstdio>
__global__ void k_step(int* counter) { if (threadIdx.x == 0 && blockIdx.x == 0) ++(*counter); }
__global__ void k_cond(const int* counter, int limit, cudaGraphConditionalHandle h) {
    if (threadIdx.x == 0) cudaGraphSetConditional(h, *counter < limit ? 1u : 0u);
}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("FAIL %s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main() {
    cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    int* d; CK(cudaMalloc(&d, 4));
    cudaGraph_t g; CK(cudaGraphCreate(&g, 0));
    int limit = 7;
    void* a1[] = {&d};
    cudaKernelNodeParams kp{}; kp.func = (void*)k_step; kp.gridDim = dim3(4); kp.blockDim = dim3(64); kp.kernelParams = a1;
    cudaGraphNode_t n1, n2, nw, b1, b2;
    CK(cudaGraphAddKernelNode(&n1, g, nullptr, 0, &kp));
    cudaGraphConditionalHandle h; CK(cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));
    void* a2[] = {&d, &limit, &h};
    cudaKernelNodeParams kc{}; kc.func = (void*)k_cond; kc.gridDim = dim3(1); kc.blockDim = dim3(32); kc.kernelParams = a2;
    CK(cudaGraphAddKernelNode(&n2, g, &n1, 1, &kc));
    cudaGraphNodeParams cp{}; cp.type = cudaGraphNodeTypeConditional; cp.conditional.handle = h; cp.conditional.type = cudaGraphCondTypeWhile; cp.conditional.size = 1;
    CK(cudaGraphAddNode(&nw, g, &n2, 1, &cp));
    cudaGraph_t body = cp.conditional.phGraph_out[0];
    CK(cudaGraphAddKernelNode(&b1, body, nullptr, 0, &kp));
    CK(cudaGraphAddKernelNode(&b2, body, &b1, 1, &kc));
    cudaGraphExec_t ex; CK(cudaGraphInstantiate(&ex, g, 0));
    int ok = 1;
    for (int lim : {7, 1, 3}) {      // the limit is baked into the node arguments: 7 every time; counter restarts per launch
        (void)lim;
        CK(cudaMemsetAsync(d, 0, 4, st)); CK(cudaGraphLaunch(ex, st)); CK(cudaStreamSynchronize(st));
        int hc; CK(cudaMemcpy(&hc, d, 4, cudaMemcpyDeviceToHost));
        printf("canary: counter %d (want 7)\n", hc);
        ok &= hc == 7;
    }
    printf(ok ? "CANARY OK\n" : "CANARY FAIL\n");
    return ok ? 0 : 1;
}
