// K3: membrane constitutive forces, one CTA per cell, positions staged in shared memory, no atomics, deterministic.
// Every term is a GATHER: per mesh element first (triangle, edge, vertex patch), then per vertex over its ring.
// Replaces RbcHighOrderModel::ParticleMechanics (reference mechanics/rbcHighOrderModel.cpp:38-207)
// and PltSimpleModel::ParticleMechanics (mechanics/pltSimpleModel.cpp:44-208) behind
// HemoCellParticleField::applyConstitutiveModel (core/hemoCellParticleField.cpp:633-675).
#include "ctx.cuh"
#include <cstdlib>
#include <cfloat>

namespace {

struct MechArgs {
  CellTypeDev t;
  int64_t first_cell, first_particle;
  const uint8_t* alive;
  const double *x, *y, *z, *vx, *vy, *vz;
  double *fx, *fy, *fz;
  double* comp[6][3];
};

struct V3 { double x, y, z; };
__device__ __forceinline__ V3 sub(V3 a, V3 b) { return {a.x-b.x, a.y-b.y, a.z-b.z}; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return {a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x}; }
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x*b.x + a.y*b.y + a.z*b.z; }
__device__ __forceinline__ double norm(V3 a) { return sqrt(a.x*a.x + a.y*a.y + a.z*a.z); }
// staged vectors are [V][3]: one address per vertex, and a stride of three 8-byte words spreads consecutive or random
// vertices over all sixteen bank pairs
__device__ __forceinline__ V3 ldv(const double* X, int, int i) { return {X[3*i], X[3*i+1], X[3*i+2]}; }
// helper/array.h:271-285
__device__ __forceinline__ void tri_area_normal(V3 v0, V3 v1, V3 v2, double& area, V3& n) {
  n = cross(sub(v1, v0), sub(v2, v0));
  const double nn = norm(n);
  // one reciprocal shared by the three components (1 ulp from the reference's three divisions; nothing downstream cancels it)
  if (nn != 0.0) { area = 0.5*nn; const double inv = 1.0/nn; n.x *= inv; n.y *= inv; n.z *= inv; }
  else { area = 0.0; n = {0.0, 0.0, 0.0}; }
}

// MODEL 0 = RbcHighOrderModel, 1 = PltSimpleModel; VISC: membrane viscosity term evaluated
//
// RBC, in four steps between three barriers:
//  1. per triangle: the signed-volume term and the area-force magnitude (AFM);
//  2. lane 0 sums the volume terms sequentially while the other warps evaluate, per edge, the link (and viscous) force
//     divided by the edge length (EF, EG) and, per vertex, the bending force of its patch (BF);
//  3. per vertex, one walk around its ring: slot j brings the neighbour r_j, the edge (v, r_j) and the triangle
//     (v, r_j, r_j+1) from one packed table word, and the vertex only scales difference vectors it rebuilds from the
//     staged positions: AFM * (centroid - x_v), EF * (r_j - x_v), BF[r_j] / n, the volume force along the raw cross
//     product (area * unit normal == cross / 2).
// Everything with a square root or a division in it is thus evaluated once per mesh element (20 k division-class
// operations per cell; the first version of the gather re-evaluated a triangle three times and an edge twice, 45 k), and
// every position is loaded once per ring walk (54 shared-memory loads per vertex against 108: after the divisions had
// gone the kernel was bound by shared-memory wavefronts, profiles/README.md r2x).  The signed-volume term and the
// centroid are formed exactly as the reference forms them (no contraction, the reference's operand order: both cancel
// five to seven digits); the other products differ from the reference's by an ulp or two in places where nothing cancels
// afterwards, and the four force families reach a vertex in ring order instead of element order
// (tests/test_gpu_parity.py::test_mechanics_parity holds 1e-12 on the total and on each family).
template <int MODEL, bool VISC, bool COMP, int NT>
__global__ void __launch_bounds__(NT, MODEL == 0 ? 3 : 1)
k_mechanics(MechArgs a) {
  extern __shared__ double sm[];
  const CellTypeDev& t = a.t;
  const int V = t.V, T = t.T, E = t.E;
  const int64_t cell = a.first_cell + blockIdx.x;
  if (!a.alive[cell]) return;
  const int64_t base = a.first_particle + (int64_t)blockIdx.x*V;
  constexpr bool RBC = (MODEL == 0);
  double* X = sm;                    // [3V]
  double* VEL = X + 3*V;             // [3V] if VISC
  double* VT = VEL + (VISC ? 3*V : 0);    // [T]  signed-volume terms
  double* TA = VT + T;               // [T+1] PLT: area          RBC: AFM, area-force magnitude (+ the null triangle's 0)
  double* TN = TA + T + 1;           // [3T]  PLT: unit normal
  double* BF = TN + (RBC ? 0 : 3*T); // [3V]  RBC: bending force of every vertex's own patch
  double* EF = BF + (RBC ? 3*V : 0); // [E+1] RBC: link force / edge length (+ the null edge's 0)
  double* EG = EF + (RBC ? E + 1 : 0);    // [E+1] RBC + VISC: viscous force / edge length
  __shared__ double s_volume;
  __shared__ double s_ninv[8];       // -1/n: the share of a patch's bending force that each of its n ring vertices takes
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid < 8) s_ninv[tid] = tid ? -1.0/(double)tid : 0.0;
  if (RBC && tid == 8) { TA[T] = 0.0; EF[E] = 0.0; if (VISC) EG[E] = 0.0; }   // the null triangle / edge of unused ring slots

  for (int i = tid; i < V; i += nt) {
    X[3*i] = a.x[base+i]; X[3*i+1] = a.y[base+i]; X[3*i+2] = a.z[base+i];
    if (VISC) { VEL[3*i] = a.vx[base+i]; VEL[3*i+1] = a.vy[base+i]; VEL[3*i+2] = a.vz[base+i]; }
  }
  __syncthreads();

  // ---- per triangle: signed-volume term (bit-exact, no contraction), area (and unit normal / area-force magnitude).
  // The vertex indices of the thread's next triangle are fetched while this one is computed.
  {
    int k = tid;
    int i0 = 0, i1 = 0, i2 = 0;
    if (k < T) { i0 = t.tri[3*k]; i1 = t.tri[3*k+1]; i2 = t.tri[3*k+2]; }
    while (k < T) {
      const int kn = k + nt;
      int n0 = 0, n1 = 0, n2 = 0;
      if (kn < T) { n0 = t.tri[3*kn]; n1 = t.tri[3*kn+1]; n2 = t.tri[3*kn+2]; }
      const double aeq = RBC ? t.tri_area_eq[k] : 0.0;
      const V3 v0 = ldv(X, V, i0), v1 = ldv(X, V, i1), v2 = ldv(X, V, i2);
      const double v210 = __dmul_rn(__dmul_rn(v2.x, v1.y), v0.z);
      const double v120 = __dmul_rn(__dmul_rn(v1.x, v2.y), v0.z);
      const double v201 = __dmul_rn(__dmul_rn(v2.x, v0.y), v1.z);
      const double v021 = __dmul_rn(__dmul_rn(v0.x, v2.y), v1.z);
      const double v102 = __dmul_rn(__dmul_rn(v1.x, v0.y), v2.z);
      const double v012 = __dmul_rn(__dmul_rn(v0.x, v1.y), v2.z);
      VT[k] = __dadd_rn(__dsub_rn(__dsub_rn(__dadd_rn(__dadd_rn(-v210, v120), v201), v021), v102), v012);
      if (RBC) {
        // rbcHighOrderModel.cpp:72-92
        const double area = 0.5*norm(cross(sub(v1, v0), sub(v2, v0)));
        const double areaRatio = (area - aeq)/aeq;
        TA[k] = t.k_area * (areaRatio + areaRatio/fabs(0.09 - areaRatio*areaRatio));
      } else {
        double area; V3 n;
        tri_area_normal(v0, v1, v2, area, n);
        TA[k] = area; TN[3*k] = n.x; TN[3*k+1] = n.y; TN[3*k+2] = n.z;
      }
      k = kn; i0 = n0; i1 = n1; i2 = n2;
    }
  }
  __syncthreads();
  // ---- the signed volume is summed sequentially in triangle order (== the reference's rounding: the absolute-coordinate
  // formula loses ~7 digits, any other order changes the volume force at 1e-10) by ONE lane: 1280 dependent additions,
  // the longest chain in the kernel (about 40 % of a cell's time on the SM).  It starts as soon as the terms exist and runs
  // beside everything that does not need the volume: the other warps' edge and bending passes.
  if (tid == 0) {
    double vol = 0.0;
    int k = 0;
    if ((reinterpret_cast<uintptr_t>(VT) & 15) == 0)
      for (; k + 1 < T; k += 2) { const double2 p = *reinterpret_cast<const double2*>(VT + k); vol = __dadd_rn(__dadd_rn(vol, p.x), p.y); }
    for (; k < T; k++) vol = __dadd_rn(vol, VT[k]);
    s_volume = __dmul_rn(vol, 1.0/6.0);
  }
  if (RBC && (tid >= 32 || nt <= 32)) {
    const int lane0 = nt > 32 ? 32 : 0, span = nt > 32 ? nt - 32 : nt;
    // ---- per edge: link force (+ membrane viscosity) over the edge length (rbcHighOrderModel.cpp:169-201)
    {
      int e = tid - lane0;
      int ia = 0, ib = 0;
      if (e < E) { ia = t.edge[2*e]; ib = t.edge[2*e+1]; }
      while (e < E) {
        const int en = e + span;
        int na = 0, nb = 0;
        if (en < E) { na = t.edge[2*en]; nb = t.edge[2*en+1]; }
        const double leq = t.edge_len_eq[e];
        const V3 ev = sub(ldv(X, V, ib), ldv(X, V, ia));
        const double len = norm(ev);
        const double ilen = 1.0/len;
        const double frac = (len - leq)/leq;
        EF[e] = (t.k_link * (frac + frac/fabs(9.0 - frac*frac)))*ilen;
        if (VISC) {
          const V3 uv = {ev.x*ilen, ev.y*ilen, ev.z*ilen};
          const V3 rv = sub(ldv(VEL, V, ib), ldv(VEL, V, ia));
          const double pr = dot(rv, uv);
          const double g0 = t.eta_m*(pr*uv.x), g1 = t.eta_m*(pr*uv.y), g2 = t.eta_m*(pr*uv.z);
          const double m2 = g0*g0 + g1*g1 + g2*g2;
          double gs = (t.eta_m*pr)*ilen;
          if (m2 > 12.5*12.5) gs *= 12.5/sqrt(m2);
          EG[e] = gs;
        }
        e = en; ia = na; ib = nb;
      }
    }
    // ---- bending force of every vertex's own patch (rbcHighOrderModel.cpp:127-158); ring vertices from the packed table
    const double inv_edge_mean = 1.0/t.edge_mean_eq;
    const uint2* rgt = reinterpret_cast<const uint2*>(t.rg);
    for (int i = tid - lane0; i < V; i += span) {
      const int nn = t.nring[i];
      const double peq = t.patch_eq[i];
      int ring[6];
#pragma unroll
      for (int j = 0; j < 6; j++) ring[j] = (int)(rgt[j*V + i].x & 0xffff);
      const V3 xi = ldv(X, V, i);
      V3 pn = {0.0, 0.0, 0.0};
      const V3 r0 = ldv(X, V, ring[0]);
      V3 prev = sub(r0, xi);
      const V3 first = prev;
      V3 sum = r0;
#pragma unroll
      for (int j = 0; j < 6; j++) {
        if (j < nn) {
          V3 nxt = first;
          if (j + 1 < nn) { const V3 r = ldv(X, V, ring[j < 5 ? j + 1 : 0]); sum.x += r.x; sum.y += r.y; sum.z += r.z; nxt = sub(r, xi); }
          const V3 tn = cross(prev, nxt);
          const double il = rsqrt(tn.x*tn.x + tn.y*tn.y + tn.z*tn.z);
          pn.x += tn.x*il; pn.y += tn.y*il; pn.z += tn.z*il;
          prev = nxt;
        }
      }
      // the ring mean minus the vertex cancels five digits: a true division, as the reference does
      const V3 mid = {sum.x/nn, sum.y/nn, sum.z/nn};
      const V3 dev = sub(mid, xi);
      const double il = rsqrt(pn.x*pn.x + pn.y*pn.y + pn.z*pn.z);
      pn.x *= il; pn.y *= il; pn.z *= il;
      const double ndev = dot(pn, dev);
      const double dDev = (ndev - peq) * inv_edge_mean;
      const double s = t.k_bend * (dDev + dDev/fabs(0.0555 - dDev*dDev));
      BF[3*i] = s*pn.x; BF[3*i+1] = s*pn.y; BF[3*i+2] = s*pn.z;
    }
  }
  __syncthreads();
  const double volume_frac = (s_volume - t.volume_eq)/t.volume_eq;
  const double volume_force = -t.k_volume * volume_frac/fabs(0.01 - volume_frac*volume_frac);

  const double third = 1.0/3.0, inv_area_mean = 1.0/t.area_mean_eq;
  if (RBC) {
    // ---- RBC, per vertex: one pass around the ring.  Slot j brings the ring vertex r_j (its position and its patch's
    // bending force), the link (v, r_j) and the triangle (v, r_j, r_j+1); every position is loaded once.  The four force
    // families are summed separately and added in the reference's order area, volume, bending, link (, viscosity);
    // within a family the terms arrive in ring order instead of element order (a re-association, ~1e-16 of the largest term).
    const double vf_half = (volume_force*0.5)*inv_area_mean;
    for (int v = tid; v < V; v += nt) {
      const V3 xv = ldv(X, V, v);
      const uint2* rgp = reinterpret_cast<const uint2*>(t.rg) + v;      // [6][V]: consecutive lanes read consecutive words
      uint2 cnext = rgp[0];
      V3 fa = {0.0, 0.0, 0.0}, fw = {0.0, 0.0, 0.0}, fl = {0.0, 0.0, 0.0}, fs = {0.0, 0.0, 0.0};
      V3 fb = ldv(BF, V, v);
      const V3 r0 = ldv(X, V, (int)(cnext.x & 0xffff));
      V3 ra = r0;
#pragma unroll 1      // unrolling the ring walk spills at the 80 registers three cells per SM allow, and measured slower
      for (int j = 0; j < 6; j++) {
        const uint2 cj = cnext;
        if (j < 5) cnext = rgp[(j + 1)*V];
        const int ia = (int)(cj.x & 0xffff), e = (int)(cj.x >> 16), tr = (int)(cj.y & 0xffff);
        const V3 rb = (j < 5) ? ldv(X, V, (int)(cnext.x & 0xffff)) : r0;
        const V3 da = sub(ra, xv), db = sub(rb, xv);
        // link (+ viscosity): the edge's scalar along (r_j - x_v)   (rbcHighOrderModel.cpp:169-201)
        const double ef = EF[e];
        fl.x += da.x*ef; fl.y += da.y*ef; fl.z += da.z*ef;
        if (VISC) { const double eg = EG[e]; fs.x += da.x*eg; fs.y += da.y*eg; fs.z += da.z*eg; }
        // bending: the reaction -F_i/n_i of the neighbour's patch (:127-158)
        const double w = s_ninv[(cj.y >> 16) & 7];
        fb.x += BF[3*ia]*w; fb.y += BF[3*ia+1]*w; fb.z += BF[3*ia+2]*w;
        // area force towards the centroid, summed (v0 + v1) + v2 as the reference does (:72-92); volume force along the
        // triangle's cross product == 2 * area * unit normal (:100-113)
        const int last = (int)((cj.y >> 19) & 3);
        const V3 p = (last == 0) ? ra : xv, q = (last == 2) ? ra : rb, l = (last == 0) ? xv : ((last == 1) ? ra : rb);
        const double afm = TA[tr];
        // (no contraction: the centroid is rounded before the vertex is subtracted, five digits cancel there)
        const double cx = __dmul_rn((p.x + q.x) + l.x, third), cy = __dmul_rn((p.y + q.y) + l.y, third), cz = __dmul_rn((p.z + q.z) + l.z, third);
        fa.x += afm*__dsub_rn(cx, xv.x); fa.y += afm*__dsub_rn(cy, xv.y); fa.z += afm*__dsub_rn(cz, xv.z);
        const V3 cr = cross(da, db);
        const double vs = ((cj.y >> 21) & 1) ? -vf_half : vf_half;
        fw.x += vs*cr.x; fw.y += vs*cr.y; fw.z += vs*cr.z;
        ra = rb;
      }
      if (COMP) {
        a.comp[0][0][base+v] = fa.x; a.comp[0][1][base+v] = fa.y; a.comp[0][2][base+v] = fa.z;
        a.comp[1][0][base+v] = fw.x; a.comp[1][1][base+v] = fw.y; a.comp[1][2][base+v] = fw.z;
        a.comp[2][0][base+v] = fb.x; a.comp[2][1][base+v] = fb.y; a.comp[2][2][base+v] = fb.z;
        a.comp[3][0][base+v] = fl.x; a.comp[3][1][base+v] = fl.y; a.comp[3][2][base+v] = fl.z;
        a.comp[4][0][base+v] = fs.x; a.comp[4][1][base+v] = fs.y; a.comp[4][2][base+v] = fs.z;
        a.comp[5][0][base+v] = 0.0; a.comp[5][1][base+v] = 0.0; a.comp[5][2][base+v] = 0.0;
      }
      double F0 = ((fa.x + fw.x) + fb.x) + fl.x, F1 = ((fa.y + fw.y) + fb.y) + fl.y, F2 = ((fa.z + fw.z) + fb.z) + fl.z;
      if (VISC) { F0 += fs.x; F1 += fs.y; F2 += fs.z; }
      a.fx[base+v] = F0; a.fy[base+v] = F1; a.fz[base+v] = F2;
    }
    return;
  }

  // ---- PLT, per vertex gather, every sum in the order the reference's loops reach the vertex
  for (int v = tid; v < V; v += nt) {
    const V3 xv = ldv(X, V, v);
    double F0 = 0.0, F1 = 0.0, F2 = 0.0;
    double c0, c1, c2;
    // area force (pltSimpleModel.cpp:73-90) and volume force (:105-117) of the incident triangles
    c0 = c1 = c2 = 0.0;
    double w0 = 0.0, w1 = 0.0, w2 = 0.0;          // volume-force sum (added after the area terms, as the reference's two loops do)
    for (int k = 0; k < 6; k++) {
      const int tr = t.vt[6*v + k]; if (tr < 0) break;
      const int i0 = t.tri[3*tr], i1 = t.tri[3*tr+1], i2 = t.tri[3*tr+2];
      const V3 v0 = ldv(X, V, i0), v1 = ldv(X, V, i1), v2 = ldv(X, V, i2);
      const double area = TA[tr];
      const double aeq = t.tri_area_eq[tr];
      const double areaRatio = (area - aeq)/aeq;
      const double afm = t.k_area * (areaRatio + areaRatio/fabs(0.09 - areaRatio*areaRatio));
      const double s = area*inv_area_mean;
      w0 += (volume_force*TN[3*tr])*s; w1 += (volume_force*TN[3*tr+1])*s; w2 += (volume_force*TN[3*tr+2])*s;
      const double cx = (v0.x+v1.x+v2.x)*third, cy = (v0.y+v1.y+v2.y)*third, cz = (v0.z+v1.z+v2.z)*third;
      const double a0 = afm*(cx - xv.x), a1 = afm*(cy - xv.y), a2 = afm*(cz - xv.z);
      F0 += a0; F1 += a1; F2 += a2;
      if (COMP) { c0 += a0; c1 += a1; c2 += a2; }
    }
    if (COMP) { a.comp[0][0][base+v] = c0; a.comp[0][1][base+v] = c1; a.comp[0][2][base+v] = c2;
                a.comp[1][0][base+v] = w0; a.comp[1][1][base+v] = w1; a.comp[1][2][base+v] = w2; c0 = c1 = c2 = 0.0; }
    F0 += w0; F1 += w1; F2 += w2;
    {
      // PLT: per edge link, viscosity, dihedral bending (pltSimpleModel.cpp:119-185)
      double l0 = 0, l1 = 0, l2 = 0, s0 = 0, s1 = 0, s2 = 0, b0 = 0, b1 = 0, b2 = 0;
      for (int k = 0; k < 12; k++) {
        const int code = t.vpe[12*v + k]; if (code < 0) break;
        const int e = code >> 2, role = code & 3;
        const int ia = t.edge[2*e], ib = t.edge[2*e+1];
        const V3 ev = sub(ldv(X, V, ib), ldv(X, V, ia));
        const double len = sqrt(ev.x*ev.x + ev.y*ev.y + ev.z*ev.z);
        const double ilen = 1.0/len;
        const V3 uv = {ev.x*ilen, ev.y*ilen, ev.z*ilen};
        if (role < 2) {
          const double sg = role ? -1.0 : 1.0;
          const double leq = t.edge_len_eq[e];
          const double frac = (len - leq)/leq;
          const double fs = t.k_link * (frac + frac/fabs(9.0 - frac*frac));
          const double a0 = uv.x*fs, a1 = uv.y*fs, a2 = uv.z*fs;
          F0 += sg*a0; F1 += sg*a1; F2 += sg*a2;
          if (COMP) { l0 += sg*a0; l1 += sg*a1; l2 += sg*a2; }
          const V3 rv = sub(ldv(VEL, V, ib), ldv(VEL, V, ia));
          const double pr = dot(rv, uv);
          double g0 = t.eta_m*(pr*uv.x), g1 = t.eta_m*(pr*uv.y), g2 = t.eta_m*(pr*uv.z);
          const double mag = sqrt(g0*g0 + g1*g1 + g2*g2);
          if (mag > 12.5) { const double s = 12.5/mag; g0 *= s; g1 *= s; g2 *= s; }
          F0 += sg*g0; F1 += sg*g1; F2 += sg*g2;
          if (COMP) { s0 += sg*g0; s1 += sg*g1; s2 += sg*g2; }
        }
        const int t0 = t.bend_tri[2*e], t1 = t.bend_tri[2*e+1];
        const V3 n1 = {TN[3*t0], TN[3*t0+1], TN[3*t0+2]}, n2 = {TN[3*t1], TN[3*t1+1], TN[3*t1+2]};
        const double angle = atan2(dot(cross(n1, n2), uv), dot(n1, n2));
        const double af = angle - t.edge_ang_eq[e];
        const double fm = t.k_bend * (af + af/fabs(2.467 - af*af));
        const double sb = (role < 2) ? 1.0 : -1.0;
        const double a0 = fm*(n1.x+n2.x)*0.5, a1 = fm*(n1.y+n2.y)*0.5, a2 = fm*(n1.z+n2.z)*0.5;
        F0 += sb*a0; F1 += sb*a1; F2 += sb*a2;
        if (COMP) { b0 += sb*a0; b1 += sb*a1; b2 += sb*a2; }
      }
      // inner links, linear (pltSimpleModel.cpp:188-205)
      double n0 = 0, n1_ = 0, n2_ = 0;
      for (int k = 0; k < 4; k++) {
        const int code = t.vin[4*v + k]; if (code < 0) break;
        const int e = code >> 1; const double sg = (code & 1) ? -1.0 : 1.0;
        const int ia = t.inner[2*e], ib = t.inner[2*e+1];
        const V3 ev = sub(ldv(X, V, ib), ldv(X, V, ia));
        const double len = sqrt(ev.x*ev.x + ev.y*ev.y + ev.z*ev.z);
        const double leq = t.inner_len_eq[e];
        const double frac = (len - leq)/leq;
        const double fs = t.k_link*5.0*frac;
        const double ilen = 1.0/len;
        const double a0 = (ev.x*ilen)*fs, a1 = (ev.y*ilen)*fs, a2 = (ev.z*ilen)*fs;
        F0 += sg*a0; F1 += sg*a1; F2 += sg*a2;
        if (COMP) { n0 += sg*a0; n1_ += sg*a1; n2_ += sg*a2; }
      }
      if (COMP) {
        a.comp[2][0][base+v] = b0; a.comp[2][1][base+v] = b1; a.comp[2][2][base+v] = b2;
        a.comp[3][0][base+v] = l0; a.comp[3][1][base+v] = l1; a.comp[3][2][base+v] = l2;
        a.comp[4][0][base+v] = s0; a.comp[4][1][base+v] = s1; a.comp[4][2][base+v] = s2;
        a.comp[5][0][base+v] = n0; a.comp[5][1][base+v] = n1_; a.comp[5][2][base+v] = n2_;
      }
    }
    a.fx[base+v] = F0; a.fy[base+v] = F1; a.fz[base+v] = F2;
  }
}

// per cell: bounding box (helper/cellInfo.cpp:170-200) ; one warp per cell
__global__ void k_bbox(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                       const int64_t* __restrict__ cell_base, const int32_t* __restrict__ cell_type,
                       const int* __restrict__ typeV, int64_t ncells, double* out) {
  const int64_t cell = (int64_t)blockIdx.x*(blockDim.x/32) + threadIdx.x/32;
  if (cell >= ncells) return;
  const int lane = threadIdx.x & 31;
  const int V = typeV[cell_type[cell]]; const int64_t b = cell_base[cell];
  double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  for (int i = lane; i < V; i += 32) {
    const double p[3] = {x[b+i], y[b+i], z[b+i]};
    for (int d = 0; d < 3; d++) { mn[d] = fmin(mn[d], p[d]); mx[d] = fmax(mx[d], p[d]); }
  }
  for (int s = 16; s > 0; s >>= 1)
    for (int d = 0; d < 3; d++) {
      mn[d] = fmin(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], s));
      mx[d] = fmax(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], s));
    }
  if (lane == 0) for (int d = 0; d < 3; d++) { out[6*cell + 2*d] = mn[d]; out[6*cell + 2*d + 1] = mx[d]; }
}

// per cell: volume and surface area (helper/cellInfo.cpp:39-101); one warp per cell, type given
__global__ void k_volume_area(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                              int64_t first_cell, int64_t first_particle, int64_t ncells, int V, int T,
                              const int* __restrict__ tri, double* vol, double* area) {
  const int64_t c = (int64_t)blockIdx.x*(blockDim.x/32) + threadIdx.x/32;
  if (c >= ncells) return;
  const int lane = threadIdx.x & 31;
  const int64_t b = first_particle + c*V;
  double v = 0.0, ar = 0.0;
  for (int k = lane; k < T; k += 32) {
    const int i0 = tri[3*k], i1 = tri[3*k+1], i2 = tri[3*k+2];
    const V3 v0 = {x[b+i0], y[b+i0], z[b+i0]}, v1 = {x[b+i1], y[b+i1], z[b+i1]}, v2 = {x[b+i2], y[b+i2], z[b+i2]};
    // translation-safe: relative to the first vertex of the cell
    const V3 o = {x[b], y[b], z[b]};
    v += dot(sub(v0, o), cross(sub(v1, o), sub(v2, o)));
    ar += 0.5*norm(cross(sub(v1, v0), sub(v2, v0)));
  }
  for (int s = 16; s > 0; s >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, s); ar += __shfl_xor_sync(0xffffffffu, ar, s); }
  if (lane == 0) { vol[first_cell + c] = v/6.0; area[first_cell + c] = ar; }
}

// CellInformationFunctionals::calculateCellStretch (helper/cellInfo.cpp:103-121): largest pairwise vertex
// distance of a cell.  One CTA per cell, positions staged in shared memory, each thread scans the pairs (i, j > i)
// of its vertices i; block-wide max through warp shuffles.
__global__ void __launch_bounds__(256)
k_stretch(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
          int64_t first_cell, int64_t first_particle, int V, double* __restrict__ out) {
  extern __shared__ double sp[];
  double* sx = sp; double* sy = sp + V; double* sz = sp + 2*V;
  __shared__ double wmax[8];
  const int64_t base = first_particle + (int64_t)blockIdx.x*V;
  for (int i = threadIdx.x; i < V; i += blockDim.x) { sx[i] = x[base + i]; sy[i] = y[base + i]; sz[i] = z[base + i]; }
  __syncthreads();
  double mx = 0.0;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    const double ax = sx[i], ay = sy[i], az = sz[i];
    for (int j = i + 1; j < V; j++) {
      const double dx = ax - sx[j], dy = ay - sy[j], dz = az - sz[j];
      mx = fmax(mx, dx*dx + dy*dy + dz*dz);
    }
  }
  for (int s = 16; s > 0; s >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, s));
  if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) mx = fmax(mx, wmax[w]);
    out[first_cell + blockIdx.x] = sqrt(mx);
  }
}

template <int MODEL, bool VISC, int NT>
hcg_status launch(hcg_ctx* c, const MechArgs& a, int64_t ncells, size_t smem, int threads, bool comp) {
  if (comp) {
    CUDA_TRY(c, cudaFuncSetAttribute(k_mechanics<MODEL, VISC, true, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_mechanics<MODEL, VISC, true, NT><<<(unsigned)ncells, threads, smem, c->stream>>>(a);
  } else {
    CUDA_TRY(c, cudaFuncSetAttribute(k_mechanics<MODEL, VISC, false, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_mechanics<MODEL, VISC, false, NT><<<(unsigned)ncells, threads, smem, c->stream>>>(a);
  }
  KERNEL_CHECK(c);
  return HCG_OK;
}

}  // namespace

hcg_status mech_apply(hcg_ctx* c, int ctype, bool components) {
  CellTypeHost& th = c->types[ctype];
  if (th.n_cells == 0) return HCG_OK;
  if (components && !c->comp_alloc) {
    for (int k = 0; k < 6; k++) for (int d = 0; d < 3; d++) {
      CUDA_TRY(c, cudaMalloc(&c->comp[k][d], sizeof(double)*c->cap_p));
      CUDA_TRY(c, cudaMemsetAsync(c->comp[k][d], 0, sizeof(double)*c->cap_p, c->stream));
    }
    c->comp_alloc = true;
  }
  if (th.d.model == HCG_MODEL_HOST) return HCG_OK;        // the caller's own model: forces arrive through hcg_cells_upload
  MechArgs a;
  a.t = th.d; a.first_cell = th.first_cell; a.first_particle = th.first_particle;
  a.alive = c->cell_alive;
  a.x = c->pos[0]; a.y = c->pos[1]; a.z = c->pos[2];
  a.vx = c->vel[0]; a.vy = c->vel[1]; a.vz = c->vel[2];
  a.fx = c->frc[0]; a.fy = c->frc[1]; a.fz = c->frc[2];
  for (int k = 0; k < 6; k++) for (int d = 0; d < 3; d++) a.comp[k][d] = components ? c->comp[k][d] : nullptr;
  const int V = th.d.V, T = th.d.T;
  const bool plt = th.d.model == HCG_MODEL_PLT_SIMPLE;
  const bool visc = plt || th.d.eta_m != 0.0;
  const size_t E = th.d.E;
  const size_t smem = sizeof(double)*((size_t)3*V + (visc ? 3*V : 0) + 1 +
                                     (plt ? 5*(size_t)T : 2*(size_t)T + 3*V + E + 1 + (visc ? E + 1 : 0)));
  if (plt) return launch<1, true, 256>(c, a, th.n_cells, smem, V >= 256 ? 256 : (V >= 128 ? 128 : 64), components);
  // RBC: three cells per SM (67 kB of shared memory each)
  const int threads = V >= 256 ? 256 : (V >= 128 ? 128 : 64);
  return visc ? launch<0, true, 256>(c, a, th.n_cells, smem, threads, components) : launch<0, false, 256>(c, a, th.n_cells, smem, threads, components);
}

hcg_status mech_bbox(hcg_ctx* c, double* out_dev) {
  if (c->ncells == 0) return HCG_OK;
  if ((int)c->types.size() != c->bbox_ntypes) {            // vertex counts per type: kept on the device across calls
    std::vector<int> hv; for (auto& t : c->types) hv.push_back(t.d.V);
    if (c->bbox_typeV) { CUDA_TRY(c, cudaStreamSynchronize(c->stream)); cudaFree(c->bbox_typeV); c->bbox_typeV = nullptr; }
    CUDA_TRY(c, cudaMalloc(&c->bbox_typeV, sizeof(int)*hv.size()));
    CUDA_TRY(c, hcg_h2d(c, c->bbox_typeV, hv.data(), sizeof(int)*hv.size()));
    c->bbox_ntypes = (int)c->types.size();
  }
  k_bbox<<<(unsigned)((c->ncells + 7)/8), 256, 0, c->stream>>>(c->pos[0], c->pos[1], c->pos[2], c->cell_base,
                                                               c->cell_type, c->bbox_typeV, c->ncells, out_dev);
  KERNEL_CHECK(c);
  return HCG_OK;
}

hcg_status mech_volume_area(hcg_ctx* c, double* vol_dev, double* area_dev) {
  for (auto& th : c->types) {
    if (th.n_cells == 0) continue;
    k_volume_area<<<(unsigned)((th.n_cells + 7)/8), 256, 0, c->stream>>>(c->pos[0], c->pos[1], c->pos[2],
        th.first_cell, th.first_particle, th.n_cells, th.d.V, th.d.T, th.d.tri, vol_dev, area_dev);
    KERNEL_CHECK(c);
  }
  return HCG_OK;
}

hcg_status mech_stretch(hcg_ctx* c, double* out_dev) {
  for (auto& th : c->types) {
    if (th.n_cells == 0) continue;
    k_stretch<<<(unsigned)th.n_cells, 256, sizeof(double)*3*th.d.V, c->stream>>>(c->pos[0], c->pos[1], c->pos[2],
        th.first_cell, th.first_particle, th.d.V, out_dev);
    KERNEL_CHECK(c);
  }
  return HCG_OK;
}
