// ref_whitebox.cpp -- white-box probes into the UNMODIFIED reference TRW-S
// (cpp/trw-s) for unit-level parity tests.  TEST INFRASTRUCTURE ONLY; linked
// into oracle/_ref/libref_trws.so by oracle/Makefile.
//
//   ref_trws_ordering      MRFEnergy::SetAutomaticOrdering (ordering.cpp:7-157)
//                          on the 4-connected grid built exactly like
//                          dispmap_super.construct_neighborhood
//                          (dispmap_super.m:279-302) + trws_mex.cpp:113-115.
//   ref_trws_update_message  TypeStereo{Linear,Quadratic}::Edge::UpdateMessage
//                          (typeStereoLinear.h:329-487,
//                           typeStereoQuadratic.h:329-501) on one edge.
//
// Private solver state is reached by re-declaring `private`/`protected` as
// `public` for the reference headers only (all std headers are included first).
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <assert.h>
#include <math.h>
#include <string>
#include <sstream>
#include <stdexcept>
#include <algorithm>
#include <limits>
#include <vector>
#include <cmath>
#include "mex.h"

#define private public
#define protected public
#include "MRFEnergy.h"
#undef private
#undef protected

static void wb_err(char *msg) { throw std::runtime_error(msg ? msg : "trw-s error"); }

// grid edge terms in the caller's order (0-based node u = r + H*c)
static void grid_terms(int H, int W, std::vector<int> &t, std::vector<int> &h)
{
    // vertical: start=(r,c) r<H-1 -> finish=(r+1,c); then finish->start
    for (int c = 0; c < W; c++) for (int r = 0; r < H - 1; r++) { t.push_back(r + H * c); h.push_back(r + 1 + H * c); }
    for (int c = 0; c < W; c++) for (int r = 0; r < H - 1; r++) { t.push_back(r + 1 + H * c); h.push_back(r + H * c); }
    // horizontal: start=(r,c) c<W-1 -> finish=(r,c+1); then finish->start
    for (int c = 0; c < W - 1; c++) for (int r = 0; r < H; r++) { t.push_back(r + H * c); h.push_back(r + H * (c + 1)); }
    for (int c = 0; c < W - 1; c++) for (int r = 0; r < H; r++) { t.push_back(r + H * (c + 1)); h.push_back(r + H * c); }
}

extern "C" int ref_trws_ordering(int H, int W, int32_t *ordering_out)
{
    typedef TypeStereoLinear T;
    try {
        const int L = 1;
        MRFEnergy<T> mrf(T::GlobalSize(L), wb_err);
        int N = H * W;
        std::vector<MRFEnergy<T>::NodeId> nodes(N);
        double zero = 0;
        for (int u = 0; u < N; u++) nodes[u] = mrf.AddNode(T::LocalSize(L), T::NodeData(&zero));
        std::vector<int> t, h;
        grid_terms(H, W, t, h);
        double data[2] = {0, 0};
        int inds[2] = {0, 0};
        for (size_t p = 0; p < t.size(); p++)
            mrf.AddEdge(nodes[t[p]], nodes[h[p]], T::EdgeData(1.0, 1.0, data, inds));
        mrf.SetAutomaticOrdering();
        for (int u = 0; u < N; u++) ordering_out[u] = nodes[u]->m_ordering;
    } catch (const std::exception &e) {
        fprintf(stderr, "ref_trws_ordering: %s\n", e.what());
        return -1;
    }
    return 0;
}

// One UpdateMessage call.  stored0/stored1 are the two position vectors in
// the order trws_mex.cpp:101-111 stacks them (q first, qprim second) with
// their argsort index vectors.  swapped = Edge::m_dir.  msg is updated in
// place; returns vMin through *vmin_out.
template <class T>
static int update_message(int L, const double *Di, double *msg,
                          const double *stored0, const double *stored1,
                          const int *order0, const int *order1,
                          double alpha, double lambda, double gamma,
                          int dir, int swapped, double *vmin_out)
{
    typedef typename T::Edge Edge;
    typedef typename T::Vector Vector;
    std::vector<double> stacked(2 * L);
    std::vector<int> inds(2 * L);
    memcpy(&stacked[0], stored0, sizeof(double) * L);
    memcpy(&stacked[L], stored1, sizeof(double) * L);
    memcpy(&inds[0], order0, sizeof(int) * L);
    memcpy(&inds[L], order1, sizeof(int) * L);
    typename T::GlobalSize KG(L);
    typename T::LocalSize KL(L);
    typename T::EdgeData ed(lambda, alpha, &stacked[0], &inds[0]);
    int sz = Edge::GetSizeInBytes(KG, KL, KL, ed);
    std::vector<char> storage(sz + 64);
    Edge *e = (Edge *)&storage[0];
    e->Initialize(KG, KL, KL, ed, NULL, NULL);
    if (swapped) e->Swap(KG, KL, KL);
    memcpy(e->GetMessagePtr()->m_data, msg, sizeof(double) * L);
    std::vector<char> buf(Edge::GetBufSizeInBytes(L * sizeof(double)) + 64);
    std::vector<double> src(Di, Di + L);
    double vmin = e->UpdateMessage(KG, KL, KL, (Vector *)&src[0], gamma, dir, &buf[0]);
    memcpy(msg, e->GetMessagePtr()->m_data, sizeof(double) * L);
    *vmin_out = vmin;
    return 0;
}

extern "C" int ref_trws_update_message(int kernel, int L, const double *Di, double *msg,
                                       const double *stored0, const double *stored1,
                                       const int *order0, const int *order1,
                                       double alpha, double lambda, double gamma,
                                       int dir, int swapped, double *vmin_out)
{
    try {
        if (kernel == 1)
            return update_message<TypeStereoLinear>(L, Di, msg, stored0, stored1, order0, order1,
                                                    alpha, lambda, gamma, dir, swapped, vmin_out);
        if (kernel == 2)
            return update_message<TypeStereoQuadratic>(L, Di, msg, stored0, stored1, order0, order1,
                                                       alpha, lambda, gamma, dir, swapped, vmin_out);
    } catch (const std::exception &e) {
        fprintf(stderr, "ref_trws_update_message: %s\n", e.what());
    }
    return -1;
}
