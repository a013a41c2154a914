// Build-side shim (test infrastructure): the reference's CPU sources include <TH/TH.h>
// only for THArgCheck (models/DCNv2/src/cpu/dcn_v2_cpu.cpp:7,145-146); TH was removed
// from torch >= 1.11. Map the macro onto TORCH_CHECK so the sources compile unmodified.
#pragma once
#include <c10/util/Exception.h>
#define THArgCheck(cond, argN, ...) TORCH_CHECK(cond, __VA_ARGS__)
