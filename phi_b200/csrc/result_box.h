// The allocation behind a phi_index_result (internal).
#pragma once
#include "../../include/phi_gpu_index.h"
#include <cstddef>

// pinned host buffer recycled through the ctx's pool (result arrays land here: D2H at full PCIe speed, no staging copy)
struct PinnedBuf { void *p = nullptr; size_t cap = 0; };

// A result owns pinned buffers borrowed from its ctx's pool; freeing it hands them back (or releases them if the ctx is gone).
// Merged results (phi_index_result_merge) own plain heap arrays instead (heap != 0, owner == NULL).
struct ResultBox {
    phi_index_result pub;              // must stay the first member: the public pointer is &box->pub
    phi_gpu_index_ctx *owner;
    PinnedBuf bufs[12]; int nbufs;
    int heap;
};
