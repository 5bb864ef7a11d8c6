// Host side of the path after the kernels: the redundancy filter of align_contigs and the final
// (aread, abpos, bread, comp) ordering.  Stays host C++ on purpose ("stays host C; must
// be bit-compatible"): it is <1 % of the work, is driven by libc qsort's tie order, and works on
// the few hundred thousand records the device returns.
//
// Replaces (reference file:line):
//   entwine                         FastGA.c:2818-2941
//   redundancy filter               FastGA.c:3407-3685   (per (A-contig, B-contig, strand) call)
//   la_sort / SORT_MAP order        FastGA.c:3800-3900
#include "common.cuh"
#include "handles.h"
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <algorithm>

#define OUT_HDR   40
#define TSPACE    100
#define BOX_FUZZ  10

struct fgb_overlaps;
extern "C" long long fgb_overlaps_bytes(const fgb_overlaps *o);
extern "C" const unsigned char *fgb_overlaps_data(const fgb_overlaps *o);

struct HPath { int abpos, bbpos, aepos, bepos, diffs, tlen; };

struct HOvl
{ HPath p;
  unsigned flags;                 // ELIMINATED / OWNS_MEMORY
  const unsigned char *trace;     // into the record buffer, or an owned fused trace
  unsigned char *owned;
};

#define ELIMINATED  0x4
#define OWNS_MEMORY 0x8

static int ALIGN_SORT(const void *l, const void *r)          // FastGA.c:2962-2967
{ const HOvl *ol = *((HOvl * const *) l), *orr = *((HOvl * const *) r);
  return (ol->p.abpos - orr->p.abpos);
}

//  Walks two trace-point paths over their common A-interval; returns the signed minimum
//  B-separation (0 if they cross) and the last trace point at which they coincide.

static int entwine(const HPath *jp, const unsigned char *jt, const HPath *kp, const unsigned char *kt,
                   int *where)
{ int ac, b2, y2, yp, ae, i, j, k, mn;

  *where = -1;
  y2 = jp->bbpos;
  b2 = kp->bbpos;
  j  = jp->abpos/TSPACE;
  k  = kp->abpos/TSPACE;
  ac = k*TSPACE;
  j = 1 + 2*(k-j);
  k = 1;
  for (i = 1; i < j; i += 2)
    y2 += jt[i];
  if (j == 1)
    yp = y2 + (jt[j] * (kp->abpos - jp->abpos)) / (ac+TSPACE - jp->abpos);
  else
    yp = y2 + (jt[j] * (kp->abpos - ac)) / TSPACE;
  mn = b2-yp;

  ae = jp->aepos;
  if (ae > kp->aepos) ae = kp->aepos;

  for (ac += TSPACE; ac < ae; ac += TSPACE)
    { y2 += jt[j];
      b2 += kt[k];
      j += 2;
      k += 2;
      i = b2-y2;
      if (mn < 0 && mn < i)      mn = (i >= 0) ? 0 : i;
      else if (mn > 0 && mn > i) mn = (i <= 0) ? 0 : i;
      if (i == 0) *where = ac;
    }

  ac -= TSPACE;
  if (ae == jp->aepos)
    { y2 = jp->bepos;
      if (kp->aepos >= ac) b2 += (kt[k] * (ae - ac)) / TSPACE;
      else                 b2 += (kt[k] * (ae - ac)) / (kp->aepos - ac);
    }
  else
    { b2 = kp->bepos;
      if (jp->aepos >= ac) y2 += (jt[j] * (ae - ac)) / TSPACE;
      else                 y2 += (jt[j] * (ae - ac)) / (jp->aepos - ac);
    }
  i = b2-y2;
  if (mn < 0 && mn < i)      mn = (i >= 0) ? 0 : i;
  else if (mn > 0 && mn > i) mn = (i <= 0) ? 0 : i;
  return mn;
}

static void filter_group(std::vector<HOvl> &g)
{ int nlas = (int) g.size(), j, k, where, dist;
  std::vector<HOvl *> perm(nlas);
  for (j = 0; j < nlas; j++) perm[j] = &g[j];
  qsort(perm.data(),nlas,sizeof(HOvl *),ALIGN_SORT);          // libc, as FastGA.c:3435

  for (j = nlas-1; j >= 0; j--)                               // pass 1 (:3441-3491)
    { HOvl *o = perm[j]; HPath *op = &o->p;
      for (k = j+1; k < nlas; k++)
        { HOvl *w = perm[k]; HPath *wp = &w->p;
          if (op->aepos <= wp->abpos) break;
          if (w->flags & ELIMINATED) continue;
          if (op->abpos == wp->abpos && op->bbpos == wp->bbpos)
            { if (op->aepos == wp->aepos && op->bepos == wp->bepos)
                { if (op->diffs < wp->aepos) { w->flags |= ELIMINATED; continue; }   // sic (:3456)
                  else                       { o->flags |= ELIMINATED; break; }
                }
              else
                { if (op->aepos > wp->aepos) { w->flags |= ELIMINATED; continue; }
                  else                       { o->flags |= ELIMINATED; break; }
                }
            }
          else if (op->aepos == wp->aepos && op->bepos == wp->bepos)
            { if (op->abpos < wp->abpos) { w->flags |= ELIMINATED; continue; }
              else                       { o->flags |= ELIMINATED; break; }
            }
        }
    }

  for (j = nlas-1; j >= 0; j--)                               // pass 2 (:3493-3592)
    { HOvl *o = perm[j]; HPath *op = &o->p;
      if (o->flags & ELIMINATED) continue;
      for (k = j+1; k < nlas; k++)
        { HOvl *w = perm[k]; HPath *wp = &w->p;
          if (op->aepos <= wp->abpos) break;
          if (w->flags & ELIMINATED) continue;
          if (op->bepos <= wp->bbpos || op->bbpos >= wp->bepos) continue;

          const unsigned char *otrace = o->trace, *wtrace = w->trace;
          dist = entwine(op,otrace,wp,wtrace,&where);
          if (where != -1)                                    // fuse o[..where] + w[where..]
            { int ocut = 2 * (((where-op->abpos)-1)/TSPACE+1);
              int wcut = 2 * (((where-wp->abpos)-1)/TSPACE+1);
              int ntlen = ocut + (wp->tlen-wcut), d = 0, h = 0, q;
              unsigned char *nt = (unsigned char *) malloc(ntlen > 0 ? ntlen : 1);
              for (q = 0; q < ocut; q += 2)
                { d += (nt[h] = otrace[q]); nt[h+1] = otrace[q+1]; h += 2; }
              for (q = wcut; q < wp->tlen; q += 2)
                { d += (nt[h] = wtrace[q]); nt[h+1] = wtrace[q+1]; h += 2; }
              if (o->flags & OWNS_MEMORY) free(o->owned);
              if (w->flags & OWNS_MEMORY) { free(w->owned); w->owned = NULL; }
              op->tlen  = ntlen;
              op->diffs = d;
              op->aepos = wp->aepos;
              op->bepos = wp->bepos;
              w->flags |= ELIMINATED;
              o->flags |= OWNS_MEMORY;
              o->owned = nt; o->trace = nt;
              continue;
            }
          if (dist != 0)                                      // BOX_ELIM (:3571-3588)
            { if ((op->aepos - op->abpos) + BOX_FUZZ >= wp->aepos - wp->abpos)
                { if (wp->aepos <= op->aepos+BOX_FUZZ && wp->bbpos >= op->bbpos-BOX_FUZZ &&
                      wp->bepos <= op->bepos+BOX_FUZZ)
                    { w->flags |= ELIMINATED; continue; }
                }
              else
                { if (op->aepos <= wp->aepos+BOX_FUZZ && op->bbpos >= wp->bbpos-BOX_FUZZ &&
                      op->bepos <= wp->bepos+BOX_FUZZ && op->abpos >= wp->abpos-BOX_FUZZ)
                    { o->flags |= ELIMINATED; continue; }
                }
            }
        }
    }

  //  survivors in perm (abpos) order, as written to the per-thread file (:3649-3680)
  std::vector<HOvl> out;
  for (j = 0; j < nlas; j++)
    if (!(perm[j]->flags & ELIMINATED))
      out.push_back(*perm[j]);
    else if (perm[j]->flags & OWNS_MEMORY)
      { free(perm[j]->owned); perm[j]->owned = NULL; }
  g.swap(out);
}

struct fgb_alns
{ long long n = 0, nraw = 0;
  std::vector<int> fields;                  // n x 9: comp aread bread abpos bbpos aepos bepos diffs tlen
  std::vector<long long> toff;
  std::vector<unsigned char> pool;
};

struct RawRef { int triple, seq; long long off; };

//  Raw device records -> discovery order -> per contig-pair filter -> final SORT_MAP order.
//  perm1/perm2 map contig ranks (seed records) to original contig numbers (FastGA.c:3007-3008).

extern "C" int fgb_filter(const fgb_overlaps *O, const int *perm1, const int *perm2, int jc_bits,
                          int ic_bits, int do_filter, fgb_alns **out)
{ long long nb = fgb_overlaps_bytes(O);
  const unsigned char *buf = fgb_overlaps_data(O);
  std::vector<RawRef> refs;
  for (long long off = 0; off < nb; )
    { const int *h = (const int *) (buf + off);
      RawRef r; r.triple = h[0]; r.seq = h[1]; r.off = off;
      refs.push_back(r);
      off += OUT_HDR + ((h[8] + 7) & ~7);
    }
  std::sort(refs.begin(),refs.end(),[](const RawRef &a, const RawRef &b)
            { return a.triple != b.triple ? a.triple < b.triple : a.seq < b.seq; });

  fgb_alns *R = new fgb_alns();
  R->nraw = (long long) refs.size();
  struct Fin { int comp, aread, bread; HOvl o; };
  std::vector<Fin> fin;
  size_t i = 0;
  while (i < refs.size())
    { int pk = ((const int *) (buf + refs[i].off))[2];
      std::vector<HOvl> g;
      size_t e = i;
      while (e < refs.size() && ((const int *) (buf + refs[e].off))[2] == pk)
        { const int *h = (const int *) (buf + refs[e].off);
          HOvl o;
          o.p.abpos = h[3]; o.p.bbpos = h[4]; o.p.aepos = h[5]; o.p.bepos = h[6];
          o.p.diffs = h[7]; o.p.tlen = h[8];
          o.flags = 0; o.owned = NULL;
          o.trace = buf + refs[e].off + OUT_HDR;
          g.push_back(o);
          e += 1;
        }
      if (do_filter) filter_group(g);
      int jc = pk & ((1 << jc_bits) - 1), ic = (pk >> jc_bits) & ((1 << ic_bits) - 1);
      int comp = (pk >> (jc_bits + ic_bits)) & 1;
      for (size_t q = 0; q < g.size(); q++)
        { Fin f; f.comp = comp; f.aread = perm1[ic]; f.bread = perm2[jc]; f.o = g[q];
          fin.push_back(f);
        }
      i = e;
    }

  //  SORT_MAP (FastGA.c:3800-3836): (aread, abpos, bread, comp); remaining ties by address in
  //  the reference (= order of arrival in its per-thread file) -> here stable order of arrival.
  std::vector<int> ord(fin.size());
  for (size_t q = 0; q < fin.size(); q++) ord[q] = (int) q;
  std::stable_sort(ord.begin(),ord.end(),[&](int a, int b)
    { const Fin &x = fin[a], &y = fin[b];
      if (x.aread != y.aread) return x.aread < y.aread;
      if (x.o.p.abpos != y.o.p.abpos) return x.o.p.abpos < y.o.p.abpos;
      if (x.bread != y.bread) return x.bread < y.bread;
      return x.comp < y.comp;
    });
  R->n = (long long) fin.size();
  R->fields.resize(R->n * 9);
  R->toff.resize(R->n);
  for (size_t q = 0; q < fin.size(); q++)
    { const Fin &f = fin[ord[q]];
      int *d = &R->fields[q*9];
      d[0] = f.comp; d[1] = f.aread; d[2] = f.bread;
      d[3] = f.o.p.abpos; d[4] = f.o.p.bbpos; d[5] = f.o.p.aepos; d[6] = f.o.p.bepos;
      d[7] = f.o.p.diffs; d[8] = f.o.p.tlen;
      R->toff[q] = (long long) R->pool.size();
      R->pool.insert(R->pool.end(),f.o.trace,f.o.trace + f.o.p.tlen);
    }
  for (size_t q = 0; q < fin.size(); q++)
    if (fin[q].o.flags & OWNS_MEMORY) free(fin[q].o.owned);
  *out = R;
  return FGB_OK;
}

extern "C" void fgb_alns_free(fgb_alns *a) { delete a; }
extern "C" long long fgb_alns_count(const fgb_alns *a) { return a->n; }
extern "C" long long fgb_alns_raw_count(const fgb_alns *a) { return a->nraw; }
extern "C" long long fgb_alns_pool_bytes(const fgb_alns *a) { return (long long) a->pool.size(); }
extern "C" int fgb_alns_get(const fgb_alns *a, int *fields /* n x 9 */, long long *toff,
                            unsigned char *pool)
{ memcpy(fields,a->fields.data(),sizeof(int)*a->fields.size());
  memcpy(toff,a->toff.data(),sizeof(long long)*a->toff.size());
  if (!a->pool.empty()) memcpy(pool,a->pool.data(),a->pool.size());
  return FGB_OK;
}
