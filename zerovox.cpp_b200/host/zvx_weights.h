// zvx_weights.h -- describe the tensors of a ggml weight context to the C ABI (zvx_tensor_desc list).
// Shared by the class mirrors (zerovox_b200.cpp) and the model wrapper (zerovox_model_b200.cpp).
#pragma once

#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "ggml.h"
#include "ggml-backend.h"

#include "../../include/zvx.h"

namespace ZeroVOX
{
    // Walk the weight context and describe every tensor whose name starts with one of the
    // prefixes.  On the CPU backend tensor->data is a host pointer (zerovox.cpp:86-91); for any
    // other buffer type the bytes are fetched with ggml_backend_tensor_get.
    struct WeightSet {
        std::vector<zvx_tensor_desc> descs;
        std::vector<std::vector<uint8_t>> staged;
    };
    inline void collect(ggml_context *ctx, const std::vector<std::string> &prefixes, WeightSet &ws)
    {
        for (ggml_tensor *t = ggml_get_first_tensor(ctx); t; t = ggml_get_next_tensor(ctx, t)) {
            const std::string name = ggml_get_name(t);
            bool want = false;
            for (const std::string &p : prefixes) want = want || name.compare(0, p.size(), p) == 0;
            if (!want) continue;
            if (t->type != GGML_TYPE_F32 && t->type != GGML_TYPE_F16)
                throw std::runtime_error("tensor '" + name + "': only F32 / F16 weights are supported");
            zvx_tensor_desc d;
            memset(&d, 0, sizeof d);
            d.name = ggml_get_name(t);
            d.dtype = t->type == GGML_TYPE_F16 ? ZVX_F16 : ZVX_F32;
            d.n_dims = ggml_n_dims(t);
            for (int i = 0; i < 4; ++i) d.ne[i] = t->ne[i];
            if (t->buffer && !ggml_backend_buffer_is_host(t->buffer)) {
                ws.staged.emplace_back(ggml_nbytes(t));
                ggml_backend_tensor_get(t, ws.staged.back().data(), 0, ggml_nbytes(t));
                d.data = ws.staged.back().data();
            } else {
                d.data = t->data;
            }
            if (!d.data) throw std::runtime_error("tensor '" + name + "' has no data (weights not loaded yet?)");
            ws.descs.push_back(d);
        }
    }

}
