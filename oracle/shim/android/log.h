// TEST INFRASTRUCTURE ONLY. The reference's logging macros (global.h:39-59)
// are printf-style only under ANDROID; this maps them to stderr (quiet unless
// SVO_REF_VERBOSE is set in the environment).
#pragma once
#include <cstdio>
#include <cstdlib>
enum { ANDROID_LOG_DEBUG, ANDROID_LOG_INFO, ANDROID_LOG_WARN, ANDROID_LOG_ERROR };
static inline bool svo_ref_verbose_() { static int v = -1; if (v < 0) v = getenv("SVO_REF_VERBOSE") ? 1 : 0; return v == 1; }
#define __android_log_print(lvl, tag, ...) \
  do { if (svo_ref_verbose_()) { fprintf(stderr, __VA_ARGS__); fputc('\n', stderr); } } while (0)
