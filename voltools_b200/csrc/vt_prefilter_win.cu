// vt_prefilter_win.cu -- windowed prefilter kernels (placeholder: forwards to the sequential variant).
#include "vt_common.cuh"
int vt_prefilter_seq(float *d_vol, int d0, int d1, int d2, cudaStream_t st);
int vt_prefilter_win(float *d_vol, int d0, int d1, int d2, cudaStream_t st) { return vt_prefilter_seq(d_vol, d0, d1, d2, st); }
