// Walker/Vose alias tables for discrete sampling.  Same interface and the same pairing
// order as the reference (src/math/AliasTable.h:12-56; strided variant
// src/core/EnvironmentMap.cpp:66-114) so that tables are identical entry for entry: the
// device picks `u.y < prob[i] ? i : alias[i]` (light.glsl:68-72,181-205).
#pragma once
#include <cstdint>
#include <utility>
#include <vector>

namespace zillum {

class AliasTable {
public:
    // `residualProbOne`: entries left over when one work list empties get probability 1
    // (the environment-map variant) instead of their scaled probability (the light variant).
    // Returns the sum of the input weights.
    template <typename T>
    static float buildStrided(T* alias, float* prob, int n, int stride, bool residualProbOne) {
        struct Item { int id; float p; };
        float sum = 0.0f;
        for (int i = 0; i < n; i++) sum += prob[(size_t)i * stride];
        const float scale = (float)n / sum;
        std::vector<Item> small, large;
        small.reserve(n); large.reserve(n);
        for (int i = 0; i < n; i++) {
            float& p = prob[(size_t)i * stride];
            p *= scale;
            (p < 1.0f ? small : large).push_back({i, p});
        }
        while (!large.empty() && !small.empty()) {
            Item s = small.back(); small.pop_back();
            Item g = large.back(); large.pop_back();
            alias[(size_t)s.id * stride] = (T)g.id;
            prob[(size_t)s.id * stride] = s.p;
            g.p += s.p - 1.0f;
            (g.p < 1.0f ? small : large).push_back(g);
        }
        for (auto* rest : {&large, &small})
            while (!rest->empty()) {
                Item e = rest->back(); rest->pop_back();
                alias[(size_t)e.id * stride] = (T)e.id;
                prob[(size_t)e.id * stride] = residualProbOne ? 1.0f : e.p;
            }
        return sum;
    }

    template <typename T>
    static std::pair<std::vector<T>, std::vector<float>> build(const std::vector<float>& pdf) {
        std::vector<T> alias(pdf.size());
        std::vector<float> prob = pdf;
        if (!pdf.empty()) buildStrided<T>(alias.data(), prob.data(), (int)pdf.size(), 1, false);
        return {alias, prob};
    }
};

}  // namespace zillum
