// placeholder until the tcgen05 attention lands (next commit): reports "unsupported" so dispatch uses the fp32-math kernel
#include "bpm_common.cuh"
int bpm_xattn_tc_supported(const bpm_attn_t* a) { (void)a; return 0; }
int bpm_xattn_fwd_tc(const bpm_attn_t*, const void*, const void*, const void*, void*, float*, cudaStream_t) { bpm_set_error("xattn_tc: not built"); return BPM_EINVAL; }
int bpm_xattn_bwd_tc(const bpm_attn_t*, const void*, const void*, const void*, const void*, const void*, const float*, float*, void*, float, void*, void*, cudaStream_t) { bpm_set_error("xattn_tc: not built"); return BPM_EINVAL; }
