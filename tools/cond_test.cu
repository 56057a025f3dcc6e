#include <cuda_runtime.h>
#include <cstdio>
__global__ void setk(cudaGraphConditionalHandle h, const int *flag) { if (threadIdx.x == 0) cudaGraphSetConditional(h, *flag ? 1u : 0u); }
__global__ void body(int *x) { x[0] += 1; }
int main() {
    cudaStream_t s; cudaStreamCreate(&s);
    int *flag, *x; cudaMalloc(&flag, 4); cudaMalloc(&x, 4); cudaMemset(x, 0, 4);
    cudaGraph_t g; cudaGraphCreate(&g, 0);
    cudaGraphConditionalHandle h; cudaGraphConditionalHandleCreate(&h, g, 0, cudaGraphCondAssignDefault);
    // capture: setk -> conditional -> (body)
    cudaStreamBeginCaptureToGraph(s, g, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
    setk<<<1, 32, 0, s>>>(h, flag);
    cudaStreamCaptureStatus st; const cudaGraphNode_t *deps; size_t nd;
    cudaStreamGetCaptureInfo(s, &st, nullptr, nullptr, &deps, &nd);
    cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
    cp.conditional.handle = h; cp.conditional.type = cudaGraphCondTypeIf; cp.conditional.size = 1;
    cudaGraphNode_t cn; cudaError_t e = cudaGraphAddNode(&cn, g, deps, nd, &cp);
    printf("addnode %s\n", cudaGetErrorString(e));
    cudaGraph_t bg = cp.conditional.phGraph_out[0];
    cudaStreamUpdateCaptureDependencies(s, &cn, 1, cudaStreamSetCaptureDependencies);
    body<<<1,1,0,s>>>(x + 0); // after conditional (always)
    cudaGraph_t g2; e = cudaStreamEndCapture(s, &g2); printf("end %s\n", cudaGetErrorString(e));
    // body graph
    cudaStream_t s2; cudaStreamCreate(&s2);
    cudaStreamBeginCaptureToGraph(s2, bg, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
    body<<<1,1,0,s2>>>(x); body<<<1,1,0,s2>>>(x);
    cudaGraph_t g3; e = cudaStreamEndCapture(s2, &g3); printf("end body %s\n", cudaGetErrorString(e));
    cudaGraphExec_t ex; e = cudaGraphInstantiate(&ex, g, 0); printf("inst %s\n", cudaGetErrorString(e));
    int one = 1, zero = 0, hx;
    cudaMemcpy(flag, &zero, 4, cudaMemcpyHostToDevice); cudaGraphLaunch(ex, s); cudaStreamSynchronize(s);
    cudaMemcpy(&hx, x, 4, cudaMemcpyDeviceToHost); printf("flag0 x=%d (want 1)\n", hx);
    cudaMemcpy(flag, &one, 4, cudaMemcpyHostToDevice); cudaGraphLaunch(ex, s); cudaStreamSynchronize(s);
    cudaMemcpy(&hx, x, 4, cudaMemcpyDeviceToHost); printf("flag1 x=%d (want 4)\n", hx);
    return 0;
}
