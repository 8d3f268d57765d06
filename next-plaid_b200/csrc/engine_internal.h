// engine_internal.h -- C++-side entry points shared between engine.cu and loader.cpp (not exported).
#pragma once
#include "../../include/plaid_b200.h"

pb_status pb_fail(pb_status s, const char *fmt, ...);
// pb_index_open without the per-token arrays; follow with pb_index_upload_tokens per chunk.
pb_status pb_index_open_begin(const pb_index_desc *d, pb_index **out);
pb_status pb_index_upload_tokens(pb_index *ix, long long tok_off, const int64_t *codes, const uint8_t *residuals,
                                 long long n, int space);
// call once after the last pb_index_upload_tokens
pb_status pb_index_finalize(pb_index *ix);
