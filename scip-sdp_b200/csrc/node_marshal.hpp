// node_marshal.hpp — host-side marshalling of branch-and-bound nodes (pure C++, no CUDA): what SCIP-SDP's sdpi.c does to a node
// before the solver sees it, for callers that hand a whole frontier of nodes to sdpcuda_solve_nodes.
//
//   * rows without an active variable are checked and dropped, rows with one active variable tighten its bounds
//     (prepareLPData, sdpi.c:1131-1290)
//   * variables with ub - lb <= epsilon are fixed: their objective goes to a constant, their matrices into the constant part
//     (sdpi.c:614-682, sdpisolver_sdpa.cpp:1015-1056)
//   * rows/columns of a block without an entry of an active variable and without a constant entry are removed, empty blocks too
//     (findEmptyRowColsSDP, sdpi.c:691-810)
//   * LP rows are split into one-sided rows (lhs first, then rhs), variable bounds follow as rows (sdpisolver_sdpa.cpp:1280-1403)
//
// The arrays produced are exactly those of scip-sdp_b200/misdp.py:Misdp.node_problem / flatten (the documented restatement, compared
// array by array in tests/test_node_marshal.py); sums are formed in the same order.  Shared by libsdpcuda and, as plain marshalling
// code without numerics, by the checker library.
#pragma once
#include "host_pool.hpp"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../include/sdpcuda.h"

namespace sdpnode {

constexpr double INF = 1e20;

struct Model
{
   int nvars = 0, nblocks = 0, maxn = 1;
   std::vector<double> obj;
   std::vector<int> blocksizes;
   // SDP entries in input order (block-major, constant part (var -1) first); var_order: entries of variables sorted by (var, block, row, col, value)
   std::vector<int> ev, eb, er, ec;
   std::vector<double> ex;
   std::vector<int> var_order;
   std::vector<int> key_order;      // all entries sorted by position (block, row, col), input order within a position
   // row nonzeros in input order and sorted by (row, variable)
   int nrows = 0;
   std::vector<int> rid, rj;
   std::vector<double> ra, lhs, rhs;
   std::vector<int> srt;
};

inline int model_build(Model& M, int nvars, const double* obj, int nblocks, const int* blocksizes, int nnz, const int* entvar, const int* entblk,
   const int* entrow, const int* entcol, const double* entval, int nrows, const int* rowbeg, const int* rowind, const double* rowval,
   const double* lhs, const double* rhs)
{
   if( nvars < 0 || nblocks < 0 || nnz < 0 || nrows < 0 ) return SDPCUDA_ERR_ARG;
   M.nvars = nvars; M.nblocks = nblocks; M.nrows = nrows;
   M.obj.assign(obj, obj + nvars);
   M.blocksizes.assign(blocksizes, blocksizes + nblocks);
   M.maxn = 1;
   for( int n : M.blocksizes ) { if( n <= 0 ) return SDPCUDA_ERR_ARG; M.maxn = std::max(M.maxn, n); }
   M.ev.assign(entvar, entvar + nnz); M.eb.assign(entblk, entblk + nnz); M.er.assign(entrow, entrow + nnz); M.ec.assign(entcol, entcol + nnz);
   M.ex.assign(entval, entval + nnz);
   for( int e = 0; e < nnz; ++e )
      if( M.ev[e] < -1 || M.ev[e] >= nvars || M.eb[e] < 0 || M.eb[e] >= nblocks || M.ec[e] < 0 || M.er[e] < M.ec[e] || M.er[e] >= M.blocksizes[M.eb[e]] )
         return SDPCUDA_ERR_ARG;
   M.var_order.clear();
   for( int e = 0; e < nnz; ++e ) if( M.ev[e] >= 0 ) M.var_order.push_back(e);
   std::stable_sort(M.var_order.begin(), M.var_order.end(), [&](int a, int b) {
      if( M.ev[a] != M.ev[b] ) return M.ev[a] < M.ev[b];
      if( M.eb[a] != M.eb[b] ) return M.eb[a] < M.eb[b];
      if( M.er[a] != M.er[b] ) return M.er[a] < M.er[b];
      if( M.ec[a] != M.ec[b] ) return M.ec[a] < M.ec[b];
      return M.ex[a] < M.ex[b]; });
   M.key_order.resize(nnz);
   std::iota(M.key_order.begin(), M.key_order.end(), 0);
   std::stable_sort(M.key_order.begin(), M.key_order.end(), [&](int a, int b) {
      if( M.eb[a] != M.eb[b] ) return M.eb[a] < M.eb[b];
      if( M.er[a] != M.er[b] ) return M.er[a] < M.er[b];
      return M.ec[a] < M.ec[b]; });
   const int rnz = nrows > 0 ? rowbeg[nrows] : 0;
   M.rid.resize(rnz); M.rj.assign(rowind, rowind + rnz); M.ra.assign(rowval, rowval + rnz);
   for( int i = 0; i < nrows; ++i )
      for( int p = rowbeg[i]; p < rowbeg[i + 1]; ++p ) { if( rowind[p] < 0 || rowind[p] >= nvars ) return SDPCUDA_ERR_ARG; M.rid[p] = i; }
   M.lhs.assign(lhs, lhs + nrows); M.rhs.assign(rhs, rhs + nrows);
   M.srt.resize(rnz);
   std::iota(M.srt.begin(), M.srt.end(), 0);
   std::stable_sort(M.srt.begin(), M.srt.end(), [&](int a, int b) { return M.rid[a] != M.rid[b] ? M.rid[a] < M.rid[b] : M.rj[a] < M.rj[b]; });
   return SDPCUDA_OK;
}

// the solver-form problem of one node (owns its arrays; view() points into them)
struct FlatNode
{
   std::vector<double> obj, entval, cval, lpval, lprhs;
   std::vector<int> blocksizes, varbeg, entblk, entrow, entcol, cblk, crow, ccol, lpbeg, lpind, active;
   double fixedobj = 0.0;
   sdpcuda_problem view() const
   {
      sdpcuda_problem p;
      memset(&p, 0, sizeof(p));
      p.m = (int)obj.size(); p.obj = obj.data(); p.nblocks = (int)blocksizes.size(); p.blocksizes = blocksizes.data();
      p.varbeg = varbeg.data(); p.entblk = entblk.data(); p.entrow = entrow.data(); p.entcol = entcol.data(); p.entval = entval.data();
      p.cnnz = (int)cval.size(); p.cblk = cblk.data(); p.crow = crow.data(); p.ccol = ccol.data(); p.cval = cval.data();
      p.nlp = (int)lprhs.size(); p.lpbeg = lpbeg.data(); p.lpind = lpind.data(); p.lpval = lpval.data(); p.lprhs = lprhs.data();
      return p;
   }
};

// Misdp.flatten(lb, ub, compress = True, skip_single_rows = True)
inline void flatten(const Model& M, const double* lb, const double* ub, double epsilon, FlatNode& F)
{
   const int nv = M.nvars;
   std::vector<char> fixed(nv);
   std::vector<int> amap(nv, -1);
   F = FlatNode();
   for( int j = 0; j < nv; ++j )
   {
      fixed[j] = (ub[j] - lb[j]) <= epsilon;
      if( !fixed[j] ) { amap[j] = (int)F.active.size(); F.active.push_back(j); }
      else F.fixedobj += M.obj[j] * lb[j];
   }
   const int m = (int)F.active.size();
   F.obj.resize(m);
   for( int k = 0; k < m; ++k ) F.obj[k] = M.obj[F.active[k]];
   // entries of the active variables
   F.varbeg.assign(m + 1, 0);
   {
      const size_t cap = M.var_order.size();            // no growth reallocations (a CLS node carries 30 k entries)
      F.entblk.reserve(cap); F.entrow.reserve(cap); F.entcol.reserve(cap); F.entval.reserve(cap);
   }
   for( int e : M.var_order )
   {
      if( fixed[M.ev[e]] ) continue;
      F.varbeg[amap[M.ev[e]] + 1]++;
      F.entblk.push_back(M.eb[e]); F.entrow.push_back(M.er[e]); F.entcol.push_back(M.ec[e]); F.entval.push_back(M.ex[e]);
   }
   for( int k = 0; k < m; ++k ) F.varbeg[k + 1] += F.varbeg[k];
   // constant part: A_0 and the fixed variables, duplicates summed in input order
   const long long mx = M.maxn;
   std::vector<long long> ckey;
   std::vector<double> csum;
   {
      // one pass over the entries in position order (sorted once per model): runs of equal position are summed in input order
      long long curkey = -1;
      double s = 0.0;
      bool have = false;
      auto flush = [&]() { if( have && s != 0.0 && std::fabs(s) > epsilon ) { ckey.push_back(curkey); csum.push_back(s); } };
      for( int e : M.key_order )
      {
         const int v = M.ev[e];
         if( v >= 0 && !fixed[v] ) continue;
         const long long key = ((long long)M.eb[e] * mx + M.er[e]) * mx + M.ec[e];
         if( !have || key != curkey ) { flush(); curkey = key; s = 0.0; have = true; }
         s += v < 0 ? M.ex[e] : -lb[v] * M.ex[e];
      }
      flush();
   }
   // rows/columns and blocks that carry nothing are removed
   std::vector<std::vector<char>> used(M.nblocks);
   for( int b = 0; b < M.nblocks; ++b ) used[b].assign(M.blocksizes[b], 0);
   for( size_t t = 0; t < F.entblk.size(); ++t ) { used[F.entblk[t]][F.entrow[t]] = 1; used[F.entblk[t]][F.entcol[t]] = 1; }
   for( long long k : ckey ) { const int b = (int)(k / (mx * mx)), r = (int)((k / mx) % mx), c = (int)(k % mx); used[b][r] = 1; used[b][c] = 1; }
   std::vector<std::vector<int>> newidx(M.nblocks);
   std::vector<int> bmap(M.nblocks, -1);
   for( int b = 0; b < M.nblocks; ++b )
   {
      newidx[b].assign(M.blocksizes[b], -1);
      int cnt = 0;
      for( int i = 0; i < M.blocksizes[b]; ++i ) if( used[b][i] ) newidx[b][i] = cnt++;
      if( cnt > 0 ) { bmap[b] = (int)F.blocksizes.size(); F.blocksizes.push_back(cnt); }
   }
   for( size_t t = 0; t < F.entblk.size(); ++t )
   {
      const int b = F.entblk[t];
      F.entrow[t] = newidx[b][F.entrow[t]]; F.entcol[t] = newidx[b][F.entcol[t]]; F.entblk[t] = bmap[b];
   }
   for( size_t t = 0; t < ckey.size(); ++t )
   {
      const long long k = ckey[t];
      const int b = (int)(k / (mx * mx)), r = (int)((k / mx) % mx), c = (int)(k % mx);
      F.cblk.push_back(bmap[b]); F.crow.push_back(newidx[b][r]); F.ccol.push_back(newidx[b][c]); F.cval.push_back(csum[t]);
   }
   // LP rows
   std::vector<double> rconst(M.nrows, 0.0);
   std::vector<int> nact(M.nrows, 0);
   for( size_t p = 0; p < M.rid.size(); ++p )
   {
      if( fixed[M.rj[p]] ) rconst[M.rid[p]] += M.ra[p] * lb[M.rj[p]];
      else if( M.ra[p] != 0.0 ) nact[M.rid[p]]++;
   }
   F.lpbeg.assign(1, 0);
   size_t q = 0;
   for( int i = 0; i < M.nrows; ++i )
   {
      const size_t q0 = q;
      while( q < M.srt.size() && M.rid[M.srt[q]] == i ) ++q;
      if( nact[i] < 2 ) continue;
      for( int side = 0; side < 2; ++side )
      {
         if( side == 0 ? !(M.lhs[i] > -INF) : !(M.rhs[i] < INF) ) continue;
         const double sg = side == 0 ? 1.0 : -1.0;
         for( size_t t = q0; t < q; ++t )
         {
            const int p = M.srt[t];
            if( fixed[M.rj[p]] || M.ra[p] == 0.0 ) continue;
            F.lpind.push_back(amap[M.rj[p]]); F.lpval.push_back(M.ra[p] * sg);
         }
         F.lpbeg.push_back((int)F.lpind.size());
         F.lprhs.push_back(side == 0 ? M.lhs[i] - rconst[i] : -(M.rhs[i] - rconst[i]));
      }
   }
   for( int k = 0; k < m; ++k )
   {
      const int j = F.active[k];
      if( lb[j] > -INF ) { F.lpind.push_back(k); F.lpval.push_back(1.0); F.lpbeg.push_back((int)F.lpind.size()); F.lprhs.push_back(lb[j]); }
      if( ub[j] < INF ) { F.lpind.push_back(k); F.lpval.push_back(-1.0); F.lpbeg.push_back((int)F.lpind.size()); F.lprhs.push_back(-ub[j]); }
   }
}

enum { NODE_SOLVE = 0, NODE_INFEASIBLE = 1, NODE_ALLFIXED = 2 };

// smallest-eigenvalue test of a small dense symmetric matrix by Cholesky of Z + shift I (lower triangle, row-major full storage)
inline bool psd_shifted(int n, std::vector<double>& Z, double shift)
{
   for( int k = 0; k < n; ++k )
   {
      double d = Z[(size_t)k * n + k] + shift;
      for( int p = 0; p < k; ++p ) d -= Z[(size_t)k * n + p] * Z[(size_t)k * n + p];
      if( !(d > 0.0) ) return false;
      const double l = std::sqrt(d);
      Z[(size_t)k * n + k] = l;
      for( int i = k + 1; i < n; ++i )
      {
         double s = Z[(size_t)i * n + k];
         for( int p = 0; p < k; ++p ) s -= Z[(size_t)i * n + p] * Z[(size_t)k * n + p];
         Z[(size_t)i * n + k] = s / l;
      }
   }
   return true;
}

// Misdp.node_problem: -> NODE_*; lbw/ubw receive the tightened bounds, F the solver-form problem (NODE_SOLVE) or only fixedobj (NODE_ALLFIXED)
inline int node_problem(const Model& M, const double* lb, const double* ub, double epsilon, double feastol, std::vector<double>& lbw,
   std::vector<double>& ubw, FlatNode& F)
{
   const int nv = M.nvars;
   lbw.assign(lb, lb + nv); ubw.assign(ub, ub + nv);
   std::vector<char> fixed(nv);
   std::vector<double> rconst(M.nrows);
   std::vector<int> nact(M.nrows);
   for( ;; )                                                           // until no bound moves (sdpi.c:3220-3225: while fixingfound)
   {
      for( int j = 0; j < nv; ++j ) if( lbw[j] > ubw[j] + epsilon ) return NODE_INFEASIBLE;
      if( M.nrows == 0 ) break;
      for( int j = 0; j < nv; ++j ) fixed[j] = (ubw[j] - lbw[j]) <= epsilon;
      std::fill(rconst.begin(), rconst.end(), 0.0); std::fill(nact.begin(), nact.end(), 0);
      for( size_t p = 0; p < M.rid.size(); ++p )
      {
         if( fixed[M.rj[p]] ) rconst[M.rid[p]] += M.ra[p] * lbw[M.rj[p]];
         else if( M.ra[p] != 0.0 ) nact[M.rid[p]]++;
      }
      for( int i = 0; i < M.nrows; ++i )
         if( nact[i] == 0 && (rconst[i] < M.lhs[i] - feastol || rconst[i] > M.rhs[i] + feastol) ) return NODE_INFEASIBLE;
      bool changed = false;
      for( size_t p = 0; p < M.rid.size(); ++p )
      {
         const int i = M.rid[p], j = M.rj[p];
         const double a = M.ra[p];
         if( fixed[j] || a == 0.0 || nact[i] != 1 ) continue;
         double lo = M.lhs[i] > -INF ? (M.lhs[i] - rconst[i]) / a : -INF;
         double hi = M.rhs[i] < INF ? (M.rhs[i] - rconst[i]) / a : INF;
         if( a < 0 )
         {
            const double lo2 = hi < INF ? hi : -INF, hi2 = lo > -INF ? lo : INF;
            lo = lo2; hi = hi2;
         }
         if( lo > lbw[j] + epsilon ) { lbw[j] = lo; changed = true; }
         if( hi < ubw[j] - epsilon ) { ubw[j] = hi; changed = true; }
      }
      if( !changed ) break;
   }
   bool allfixed = true;
   for( int j = 0; j < nv; ++j )
   {
      if( lbw[j] > ubw[j] + epsilon ) return NODE_INFEASIBLE;
      if( (ubw[j] - lbw[j]) > epsilon ) allfixed = false;
   }
   if( allfixed )
   {
      F = FlatNode();
      for( int j = 0; j < nv; ++j ) F.fixedobj += M.obj[j] * lbw[j];
      // Z(y) = sum_j y_j A_j - A_0 psd up to feastol for every block?
      for( int b = 0; b < M.nblocks; ++b )
      {
         const int n = M.blocksizes[b];
         std::vector<double> Z((size_t)n * n, 0.0);
         for( size_t e = 0; e < M.ev.size(); ++e )
         {
            if( M.eb[e] != b ) continue;
            const double v = M.ev[e] < 0 ? -M.ex[e] : lbw[M.ev[e]] * M.ex[e];
            Z[(size_t)M.er[e] * n + M.ec[e]] += v;
         }
         if( !psd_shifted(n, Z, feastol * (1.0 + 1e-6) + 1e-13) ) return NODE_INFEASIBLE;
      }
      return NODE_ALLFIXED;
   }
   flatten(M, lbw.data(), ubw.data(), epsilon, F);
   return NODE_SOLVE;
}

} // namespace sdpnode

// ---- the C ABI entry points built on the marshalling above; identical in every library that provides sdpcuda_solve_batch --------------
struct sdpcuda_model { sdpnode::Model M; };

namespace sdpnode {

inline int solve_nodes(sdpcuda_handle* h, const sdpcuda_model* model, int count, const double* lb, const double* ub, const sdpcuda_params* par,
   const double* cutoff, int* status, sdpcuda_result* res, double* bound, double* y, double* lbout, double* ubout)
{
   if( h == nullptr || model == nullptr || par == nullptr || count < 0 || (count > 0 && (lb == nullptr || ub == nullptr || status == nullptr)) )
      return SDPCUDA_ERR_ARG;
   const Model& M = model->M;
   const int nv = M.nvars;
   const double epsilon = 1e-9;
   const double feastol = par->feastol > 0 ? par->feastol : 1e-6;
   // presolve + marshalling of every node: independent host work, spread over the host threads (host_pool.hpp)
   const auto tn0 = std::chrono::steady_clock::now();
   std::vector<FlatNode> all(count);
   sdphost::Pool::get().run(count, [&](int i)
   {
      std::vector<double> lbw, ubw;
      FlatNode& F = all[i];
      status[i] = node_problem(M, lb + (size_t)i * nv, ub + (size_t)i * nv, epsilon, feastol, lbw, ubw, F);
      if( lbout != nullptr ) std::copy(lbw.begin(), lbw.end(), lbout + (size_t)i * nv);
      if( ubout != nullptr ) std::copy(ubw.begin(), ubw.end(), ubout + (size_t)i * nv);
      if( res != nullptr ) memset(&res[i], 0, sizeof(sdpcuda_result));
      if( bound != nullptr ) bound[i] = (status[i] == NODE_ALLFIXED) ? F.fixedobj : 0.0;
      if( y != nullptr && status[i] != NODE_INFEASIBLE ) std::copy(lbw.begin(), lbw.end(), y + (size_t)i * nv);   // fixed variables sit at their value
   });
   std::vector<FlatNode> flats;
   std::vector<int> owner;
   flats.reserve(count);
   for( int i = 0; i < count; ++i )
      if( status[i] == NODE_SOLVE ) { flats.push_back(std::move(all[i])); owner.push_back(i); }
   const int ns = (int)flats.size();
   if( ns == 0 ) return SDPCUDA_OK;
   std::vector<sdpcuda_problem> views(ns);
   std::vector<const sdpcuda_problem*> vp(ns);
   std::vector<std::vector<double>> ys(ns);
   std::vector<double*> yp(ns);
   std::vector<double> limits(ns, 1e20);
   std::vector<sdpcuda_result> rs(ns);
   for( int k = 0; k < ns; ++k )
   {
      views[k] = flats[k].view(); vp[k] = &views[k];
      ys[k].assign(std::max<size_t>(flats[k].obj.size(), 1), 0.0); yp[k] = ys[k].data();
      if( cutoff != nullptr && cutoff[owner[k]] < 1e20 ) limits[k] = cutoff[owner[k]] - flats[k].fixedobj;
   }
   const bool bprof = getenv("SDPCUDA_BATCH_PROFILE") != nullptr;
   const auto tb0 = std::chrono::steady_clock::now();
   if( bprof ) fprintf(stderr, "[nodes] %d nodes, %d to solve: presolve + marshalling %.2f ms\n", count, ns, 1e3 * std::chrono::duration<double>(tb0 - tn0).count());
   int rc = sdpcuda_solve_batch(h, ns, vp.data(), par, rs.data(), yp.data(), cutoff != nullptr ? limits.data() : nullptr);
   if( bprof ) fprintf(stderr, "[nodes] solve_batch %.2f ms\n", 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - tb0).count());
   if( rc != SDPCUDA_OK ) return rc;
   for( int k = 0; k < ns; ++k )
   {
      const int i = owner[k];
      if( res != nullptr ) res[i] = rs[k];
      if( bound != nullptr ) bound[i] = rs[k].dobj + flats[k].fixedobj;
      if( y != nullptr )
         for( size_t t = 0; t < flats[k].active.size(); ++t ) y[(size_t)i * nv + flats[k].active[t]] = ys[k][t];
   }
   return SDPCUDA_OK;
}

inline int debug_node_problem(const sdpcuda_model* model, const double* lb, const double* ub, double feastol, int* status, double* fixedobj,
   int* sizes, int* ibuf, size_t icap, double* dbuf, size_t dcap, double* lbout, double* ubout)
{
   if( model == nullptr || lb == nullptr || ub == nullptr || status == nullptr || fixedobj == nullptr || sizes == nullptr ) return SDPCUDA_ERR_ARG;
   FlatNode F;
   std::vector<double> lbw, ubw;
   *status = node_problem(model->M, lb, ub, 1e-9, feastol, lbw, ubw, F);
   *fixedobj = F.fixedobj;
   if( lbout != nullptr ) std::copy(lbw.begin(), lbw.end(), lbout);
   if( ubout != nullptr ) std::copy(ubw.begin(), ubw.end(), ubout);
   const size_t m = F.obj.size(), nb = F.blocksizes.size(), nnz = F.entval.size(), cn = F.cval.size(), nlp = F.lprhs.size(), lnz = F.lpval.size();
   sizes[0] = (int)m; sizes[1] = (int)nb; sizes[2] = (int)nnz; sizes[3] = (int)cn; sizes[4] = (int)nlp; sizes[5] = (int)lnz;
   if( *status != NODE_SOLVE ) return SDPCUDA_OK;
   const size_t ineed = nb + (m + 1) + 3 * nnz + 3 * cn + (nlp + 1) + lnz + m, dneed = m + nnz + cn + lnz + nlp;
   if( ibuf == nullptr || dbuf == nullptr || icap < ineed || dcap < dneed ) return SDPCUDA_OK;      // sizes only
   int* ip = ibuf;
   double* dp = dbuf;
   auto puti = [&](const std::vector<int>& v) { std::copy(v.begin(), v.end(), ip); ip += v.size(); };
   auto putd = [&](const std::vector<double>& v) { std::copy(v.begin(), v.end(), dp); dp += v.size(); };
   puti(F.blocksizes); puti(F.varbeg); puti(F.entblk); puti(F.entrow); puti(F.entcol); puti(F.cblk); puti(F.crow); puti(F.ccol);
   puti(F.lpbeg); puti(F.lpind); puti(F.active);
   putd(F.obj); putd(F.entval); putd(F.cval); putd(F.lpval); putd(F.lprhs);
   return SDPCUDA_OK;
}

} // namespace sdpnode

// the extern "C" wrappers, emitted once per library (define SDPNODE_EMIT_ABI before including this header in exactly one source file)
#ifdef SDPNODE_EMIT_ABI
extern "C" {
int sdpcuda_model_create(sdpcuda_model** model, int nvars, const double* obj, int nblocks, const int* blocksizes, int nnz, const int* entvar,
   const int* entblk, const int* entrow, const int* entcol, const double* entval, int nrows, const int* rowbeg, const int* rowind,
   const double* rowval, const double* lhs, const double* rhs)
{
   if( model == nullptr ) return SDPCUDA_ERR_ARG;
   sdpcuda_model* m = new sdpcuda_model();
   int rc = sdpnode::model_build(m->M, nvars, obj, nblocks, blocksizes, nnz, entvar, entblk, entrow, entcol, entval, nrows, rowbeg, rowind, rowval, lhs, rhs);
   if( rc != SDPCUDA_OK ) { delete m; return rc; }
   *model = m;
   return SDPCUDA_OK;
}
int sdpcuda_model_destroy(sdpcuda_model* model) { delete model; return SDPCUDA_OK; }
int sdpcuda_solve_nodes(sdpcuda_handle* h, const sdpcuda_model* model, int count, const double* lb, const double* ub, const sdpcuda_params* par,
   const double* cutoff, int* status, sdpcuda_result* res, double* bound, double* y, double* lbout, double* ubout)
{
   return sdpnode::solve_nodes(h, model, count, lb, ub, par, cutoff, status, res, bound, y, lbout, ubout);
}
int sdpcuda_debug_node_problem(const sdpcuda_model* model, const double* lb, const double* ub, double feastol, int* status, double* fixedobj,
   int* sizes, int* ibuf, size_t icap, double* dbuf, size_t dcap, double* lbout, double* ubout)
{
   return sdpnode::debug_node_problem(model, lb, ub, feastol, status, fixedobj, sizes, ibuf, icap, dbuf, dcap, lbout, ubout);
}
}
#endif
