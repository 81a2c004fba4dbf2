// The persistent frame kernel for weight-only FP8 models (CSMB_WEIGHTS_E4M3): frame_kernel.cu compiled a second time with
// one-byte weights — the same units, ring stages and schedule, e4m3 widened to fp32 in consume(), the per-output-channel scale
// applied where a row is finalised — into namespace csmb::fk_e4m3 (kernel + launcher only; the C entry points and their
// dispatch on csmb_model.weight_format are in frame_kernel.cu).
#define CSMB_FK_FMT 1
#include "frame_kernel.cu"
