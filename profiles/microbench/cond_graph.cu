#include <cuda_runtime.h>
#include <cstdio>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
__global__ void coop(int *x) { cg::grid_group g = cg::this_grid(); if (g.thread_rank() == 0) x[1] = 5; g.sync(); if (g.thread_rank() == g.size() - 1) x[0] += x[1]; }
__global__ void setter(cudaGraphConditionalHandle h, const int *flag) { cudaGraphSetConditional(h, *flag != 0); }
__global__ void body(int *x) { x[0] += 1; }
int main() {
    int *flag, *x; cudaMalloc(&flag, 4); cudaMalloc(&x, 8); cudaMemset(x, 0, 8);
    cudaStream_t s; cudaStreamCreate(&s);
    cudaGraph_t g; cudaGraphCreate(&g, 0);
    cudaGraphConditionalHandle h; cudaGraphConditionalHandleCreate(&h, g, 0, cudaGraphCondAssignDefault);
    // capture: setter kernel, then conditional node with body
    cudaStreamBeginCaptureToGraph(s, g, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
    setter<<<1,1,0,s>>>(h, flag);
    cudaStreamCaptureStatus st; const cudaGraphNode_t *deps; size_t ndeps; cudaGraph_t cg;
    cudaStreamGetCaptureInfo_v2(s, &st, nullptr, &cg, &deps, &ndeps);
    cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeConditional; p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeIf; p.conditional.size = 1;
    cudaGraphNode_t node; cudaError_t e = cudaGraphAddNode(&node, cg, deps, ndeps, &p);
    printf("add node: %s\n", cudaGetErrorString(e));
    cudaGraph_t bodyg = p.conditional.phGraph_out[0];
    cudaStreamUpdateCaptureDependencies(s, &node, 1, cudaStreamSetCaptureDependencies);
    cudaStream_t s2; cudaStreamCreate(&s2);
    cudaStreamBeginCaptureToGraph(s2, bodyg, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
    body<<<1,1,0,s2>>>(x);
    { void *args[] = {(void *)&x}; cudaError_t ec = cudaLaunchCooperativeKernel((const void *)coop, dim3(148 * 2), dim3(256), args, 0, s2); printf("coop capture: %s\n", cudaGetErrorString(ec)); }
    cudaStreamEndCapture(s2, nullptr);
    body<<<1,1,0,s>>>(x + 0);
    cudaStreamEndCapture(s, nullptr);
    cudaGraphExec_t ex; e = cudaGraphInstantiate(&ex, g, 0); printf("inst: %s\n", cudaGetErrorString(e));
    for (int f = 0; f < 2; ++f) { cudaMemcpy(flag, &f, 4, cudaMemcpyHostToDevice); cudaGraphLaunch(ex, s); cudaStreamSynchronize(s); int hx; cudaMemcpy(&hx, x, 4, cudaMemcpyDeviceToHost); printf("flag %d -> x %d\n", f, hx); }
    return 0;
}
