/* CPU ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * Plain C++ restatement of the reference's descriptor matching and DBoW2 slice:
 *   corbslam_client/src/ORBmatcher.cc:37-39 (thresholds), :162-291 SearchByBoW(KF,Frame), :294-423 SearchByBoWInServer,
 *     :657-790 SearchByBoW(KF,KF), :1746-1787 ComputeThreeMaxima, :1792-1808 DescriptorDistance
 *   corbslam_client/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1194 (transform of a feature set),
 *     :1218-1259 (descent), :1338-1424 (text loader); FORB.cpp:81-101; BowVector.cpp:34-84;
 *     FeatureVector.cpp:31-45; ScoringObject.cpp:23-68 (L1)
 *
 * PARITY STATUS: PINNED against the reference's own ORBmatcher.cc and Thirdparty/DBoW2 compiled unmodified into
 * oracle/_ref/libref.so, on the real ORBvoc.txt (k 10, L 6, 1 082 073 nodes): BowVector / FeatureVector / L1 score bits and
 * the match arrays of the three SearchByBoW variants are equal (tests/test_ref_cpu.py). Also: source constants (TH_LOW 50, TH_HIGH 100, HISTO_LENGTH 30), the SWAR popcount identity checked against
 * __builtin_popcount, and the real vocabulary header (k 10, L 6, L1_NORM, TF_IDF) where the file is available.
 *
 * Two reference behaviours that depend on uninitialised memory are given canonical definitions:
 *   - loadFromTextFile's `while(!f.eof())` appends a bogus child of the root with an uninitialised descriptor when the
 *     file ends with a newline; blank lines are ignored here.
 *   - if the descent reaches a leaf above level L-levelsup, `nid` is never written; here it is the leaf's node id.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <vector>

namespace {

const int TH_LOW = 50, HISTO_LENGTH = 30; /* ORBmatcher.cc:37-39 */

inline int hamming256(const uint8_t* a, const uint8_t* b) { /* ORBmatcher.cc:1792-1808, FORB.cpp:81-101 */
    const int32_t* pa = (const int32_t*)a;
    const int32_t* pb = (const int32_t*)b;
    int dist = 0;
    for (int i = 0; i < 8; i++, pa++, pb++) {
        unsigned int v = *pa ^ *pb;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) { /* :1746-1787 */
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

struct Fv { /* DBoW2::FeatureVector flattened: ascending node ids, CSR of feature indices */
    const uint32_t* nodes; const int32_t* off; const uint32_t* idx; int n;
};

struct VocNode {
    int parent = 0;
    std::vector<int> children;
    uint8_t desc[32];
    double weight = 0;
    int word_id = 0; /* Node() default (TemplatedVocabulary.h Node ctor) */
    bool is_word = false;
    bool leaf() const { return children.empty(); }
};

} // namespace

struct oracle_voc {
    int k, L, scoring, weighting;
    std::vector<VocNode> nodes; /* index = DBoW2 node id, 0 = root */
    int n_words = 0;
};

extern "C" {

int oracle_hamming256(const uint8_t* a, const uint8_t* b) { return hamming256(a, b); }

/* variant 0: SearchByBoW(KeyFrame*,Frame&) ; 1: SearchByBoWInServer ; 2: SearchByBoW(KeyFrame*,KeyFrame*).
 * Side A = the keyframe whose MapPoints are matched (pKF / pKF1), side B = Frame F / KeyFrame F / pKF2.
 * validA[i] = pMP && !pMP->isBad() of side A; validB likewise (variant 2 only, may be NULL otherwise).
 * match: variants 0,1 -> size nB, match[b] = index of the A feature whose MapPoint was assigned, else -1
 *        variant 2    -> size nA, match[a] = index of the matched B feature, else -1.     Returns nmatches. */
int oracle_search_by_bow(int variant, const uint8_t* descA, int nA, const uint8_t* descB, int nB,
                         const uint32_t* nodesA, const int32_t* offA, const uint32_t* idxA, int nnA,
                         const uint32_t* nodesB, const int32_t* offB, const uint32_t* idxB, int nnB,
                         const uint8_t* validA, const uint8_t* validB, const float* anglesA, const float* anglesB,
                         float nnratio, int check_ori, int32_t* match) {
    const bool kfkf = variant == 2;
    const int nOut = kfkf ? nA : nB;
    for (int i = 0; i < nOut; i++) match[i] = -1;
    std::vector<uint8_t> matchedB(nB, 0);
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    int nmatches = 0;
    int ia = 0, ib = 0;
    while (ia < nnA && ib < nnB) {
        if (nodesA[ia] == nodesB[ib]) {
            for (int i1 = offA[ia]; i1 < offA[ia + 1]; i1++) {
                const int a = (int)idxA[i1];
                if (!validA[a]) continue;
                const uint8_t* dA = descA + (size_t)a * 32;
                int best1 = 256, bestIdx = -1, best2 = 256;
                for (int i2 = offB[ib]; i2 < offB[ib + 1]; i2++) {
                    const int b = (int)idxB[i2];
                    if (matchedB[b]) continue;
                    if (kfkf && !validB[b]) continue;
                    const int dist = hamming256(dA, descB + (size_t)b * 32);
                    if (dist < best1) { best2 = best1; best1 = dist; bestIdx = b; }
                    else if (dist < best2) { best2 = dist; }
                }
                const bool near = kfkf ? best1 < TH_LOW : best1 <= TH_LOW;
                if (near && (float)best1 < nnratio * (float)best2) {
                    matchedB[bestIdx] = 1;
                    if (kfkf) match[a] = bestIdx; else match[bestIdx] = a;
                    if (check_ori) {
                        float rot = anglesA[a] - anglesB[bestIdx];
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)roundf(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        rotHist[bin].push_back(kfkf ? a : bestIdx);
                    }
                    nmatches++;
                }
            }
            ia++; ib++;
        } else if (nodesA[ia] < nodesB[ib]) {
            ia = (int)(std::lower_bound(nodesA + ia, nodesA + nnA, nodesB[ib]) - nodesA);
        } else {
            ib = (int)(std::lower_bound(nodesB + ib, nodesB + nnB, nodesA[ia]) - nodesB);
        }
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int j : rotHist[i]) { match[j] = -1; nmatches--; }
        }
    }
    return nmatches;
}

/* ---------------------------------------------------------------- vocabulary */
/* nodes 1..n in DBoW2 id order (file order): parent id, leaf flag, 32-byte descriptor, weight */
oracle_voc* oracle_voc_create(int k, int L, int scoring, int weighting, int n, const int32_t* parent, const uint8_t* is_leaf,
                              const uint8_t* desc, const double* weight) {
    oracle_voc* v = new oracle_voc;
    v->k = k; v->L = L; v->scoring = scoring; v->weighting = weighting;
    v->nodes.resize(n + 1);
    for (int i = 1; i <= n; i++) {
        VocNode& nd = v->nodes[i];
        nd.parent = parent[i - 1];
        if (nd.parent < 0 || nd.parent >= i) { delete v; return nullptr; }
        v->nodes[nd.parent].children.push_back(i);
        memcpy(nd.desc, desc + (size_t)(i - 1) * 32, 32);
        nd.weight = weight[i - 1];
        if (is_leaf[i - 1]) { nd.word_id = v->n_words++; nd.is_word = true; }
    }
    return v;
}
oracle_voc* oracle_voc_load_text(const char* path) { /* TemplatedVocabulary.h:1338-1424 */
    FILE* f = fopen(path, "r");
    if (!f) return nullptr;
    std::vector<char> line(1 << 16);
    if (!fgets(line.data(), (int)line.size(), f)) { fclose(f); return nullptr; }
    int k, L, n1, n2;
    if (sscanf(line.data(), "%d %d %d %d", &k, &L, &n1, &n2) != 4 || k < 0 || k > 20 || L < 1 || L > 10 || n1 < 0 || n1 > 5 ||
        n2 < 0 || n2 > 3) { fclose(f); return nullptr; }
    std::vector<int32_t> parent; std::vector<uint8_t> leaf, desc; std::vector<double> weight;
    while (fgets(line.data(), (int)line.size(), f)) {
        char* p = line.data();
        char* e;
        long pid = strtol(p, &e, 10);
        if (e == p) continue; /* blank line (see header) */
        p = e;
        long isleaf = strtol(p, &e, 10); p = e;
        for (int i = 0; i < 32; i++) { long b = strtol(p, &e, 10); p = e; desc.push_back((uint8_t)b); }
        double w = strtod(p, &e);
        parent.push_back((int32_t)pid); leaf.push_back(isleaf > 0); weight.push_back(w);
    }
    fclose(f);
    return oracle_voc_create(k, L, n1, n2, (int)parent.size(), parent.data(), leaf.data(), desc.data(), weight.data());
}
void oracle_voc_destroy(oracle_voc* v) { delete v; }
int oracle_voc_info(const oracle_voc* v, int* k, int* L, int* scoring, int* weighting, int* n_nodes, int* n_words) {
    *k = v->k; *L = v->L; *scoring = v->scoring; *weighting = v->weighting;
    *n_nodes = (int)v->nodes.size(); *n_words = v->n_words;
    return 0;
}
/* export nodes 1..n (for building the GPU vocabulary from the same data in tests) */
void oracle_voc_export(const oracle_voc* v, int32_t* parent, uint8_t* is_leaf, uint8_t* desc, double* weight) {
    for (size_t i = 1; i < v->nodes.size(); i++) {
        parent[i - 1] = v->nodes[i].parent;
        is_leaf[i - 1] = v->nodes[i].is_word;
        memcpy(desc + (i - 1) * 32, v->nodes[i].desc, 32);
        weight[i - 1] = v->nodes[i].weight;
    }
}

/* per-feature descent (:1218-1259): word id, weight, node id at level L - levelsup */
int oracle_voc_transform(const oracle_voc* v, const uint8_t* desc, int n, int levelsup, uint32_t* word_id, double* weight,
                         uint32_t* node_id) {
    if (v->nodes.size() < 2) return -1;
    const int nid_level = v->L - levelsup;
    for (int i = 0; i < n; i++) {
        const uint8_t* d = desc + (size_t)i * 32;
        int final_id = 0, level = 0, nid = -1;
        if (nid_level <= 0) nid = 0;
        do {
            ++level;
            const std::vector<int>& ch = v->nodes[final_id].children;
            final_id = ch[0];
            int best = hamming256(d, v->nodes[final_id].desc);
            for (size_t c = 1; c < ch.size(); c++) {
                int dd = hamming256(d, v->nodes[ch[c]].desc);
                if (dd < best) { best = dd; final_id = ch[c]; }
            }
            if (level == nid_level) nid = final_id;
        } while (!v->nodes[final_id].leaf());
        if (nid < 0) nid = final_id; /* canonical definition, see header */
        word_id[i] = (uint32_t)v->nodes[final_id].word_id;
        weight[i] = v->nodes[final_id].weight;
        node_id[i] = (uint32_t)nid;
    }
    return 0;
}

/* BowVector + FeatureVector of a feature set (:1127-1194 for TF_IDF/TF weighting with L1 normalisation, the only
 * configuration ORBvoc.txt uses: header "10 6 0 0"). out_words/out_vals sized n; fv_* sized n (+1 for off).
 * Returns the number of BowVector entries; *n_fv = number of FeatureVector nodes. */
int oracle_bow_build(int n, const uint32_t* word_id, const double* weight, const uint32_t* node_id, uint32_t* out_words,
                     double* out_vals, uint32_t* fv_nodes, int32_t* fv_off, uint32_t* fv_idx, int* n_fv) {
    std::map<uint32_t, double> bow;
    std::map<uint32_t, std::vector<uint32_t>> fv;
    for (int i = 0; i < n; i++) {
        if (weight[i] > 0) {
            auto it = bow.lower_bound(word_id[i]);           /* BowVector::addWeight */
            if (it != bow.end() && it->first == word_id[i]) it->second += weight[i];
            else bow.insert(it, {word_id[i], weight[i]});
            fv[node_id[i]].push_back((uint32_t)i);           /* FeatureVector::addFeature */
        }
    }
    double norm = 0.0;                                       /* BowVector::normalize(L1) */
    for (auto& kv : bow) norm += fabs(kv.second);
    if (norm > 0.0) for (auto& kv : bow) kv.second /= norm;
    int m = 0;
    for (auto& kv : bow) { out_words[m] = kv.first; out_vals[m] = kv.second; m++; }
    int g = 0, o = 0;
    for (auto& kv : fv) {
        fv_nodes[g] = kv.first; fv_off[g] = o;
        for (uint32_t i : kv.second) fv_idx[o++] = i;
        g++;
    }
    fv_off[g] = o;
    *n_fv = g;
    return m;
}

double oracle_bow_score_l1(int n1, const uint32_t* w1, const double* v1, int n2, const uint32_t* w2, const double* v2) {
    int i = 0, j = 0; /* ScoringObject.cpp:23-68 */
    double score = 0;
    while (i < n1 && j < n2) {
        if (w1[i] == w2[j]) { score += fabs(v1[i] - v2[j]) - fabs(v1[i]) - fabs(v2[j]); ++i; ++j; }
        else if (w1[i] < w2[j]) i = (int)(std::lower_bound(w1 + i, w1 + n1, w2[j]) - w1);
        else j = (int)(std::lower_bound(w2 + j, w2 + n2, w1[i]) - w2);
    }
    return -score / 2.0;
}

} // extern "C"
