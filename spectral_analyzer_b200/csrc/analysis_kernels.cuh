// analysis_kernels.cuh -- downconvert (NCO + polyphase FIR decimate) and Welch PSD kernels.
#pragma once
#include "decode.cuh"
