// dbx_tilekey.cuh — tile solver (dbx_tiles.cu): which class and bin a constraint falls into.  Called where a constraint's colour
// becomes final -- k_mark_solve for a colour kept from the step before, k_colour for a fresh one, k_mark_solve for the joints
// (their colours are the host's) -- so the tiled sort needs no pass of its own over the contacts.
//   class L  both dynamic bodies in one tile (or one body is not dynamic): the tile's CTA owns the row
//   class B  the bodies sit in tiles s and s + 1, the one in s in the right half of its tile and the one in s + 1 in the left
//            half: CTA s solves the row after its local ones.  The halves make the claim exclusive without any arbitration: a
//            body can only ever be reached across the boundary on its own side.
//   class G  everything else (a reach across more than one boundary or from the wrong half, a gear joint, an overflow colour)
// Bins: [class L: tile * 64 + colour | class B: (P + tile) * 64 + colour | class G: 2 P * 64 + colour].
#pragma once
#include "dbx_util.cuh"

namespace dbx {

constexpr int kRefGlobalBit = (int)0x80000000;      // = kRefGlobal of dbx_solver.cuh (a body reference that carries a body id)
DBX_D int tile_classify(const DevWorld& W, int bA, int bB, int col, bool forceGlobal, int2* bref) {
  const int P = W.nTiles, T = W.tileBodies;
  const int sA = W.b_tslot[bA], sB = W.b_tslot[bB];
  const int tA = sA >= 0 ? sA / T : -1, tB = sB >= 0 ? sB / T : -1;
  int cls, owner = 0;
  if (forceGlobal || col >= kTileColours || (tA < 0 && tB < 0)) cls = 2;
  else if (tA < 0 || tB < 0 || tA == tB) { cls = 0; owner = tA >= 0 ? tA : tB; }
  else {
    const int lo = min(tA, tB);
    const int sLo = tA < tB ? sA : sB, sHi = tA < tB ? sB : sA;            // slots of the body in the left / right tile
    const bool near = max(tA, tB) - lo == 1 && 2 * (sLo - lo * T) >= T && 2 * (sHi - (lo + 1) * T) < T;
    if (near) { cls = 1; owner = lo; } else cls = 2;
  }
  if (cls == 2) {
    if (sA >= 0) atomicOr(&W.b_xflag[sA], XF_G);
    if (sB >= 0) atomicOr(&W.b_xflag[sB], XF_G);
    *bref = make_int2(bA | kRefGlobalBit, bB | kRefGlobalBit);
    return 2 * P * kTileColours + min(col, kMaxColours - 1);
  }
  if (cls == 1) {
    atomicOr(&W.b_xflag[sA], tA == owner ? XF_OWNB : XF_FOREIGN);
    atomicOr(&W.b_xflag[sB], tB == owner ? XF_OWNB : XF_FOREIGN);
  }
  *bref = make_int2(tA == owner ? sA : (bA | kRefGlobalBit), tB == owner ? sB : (bB | kRefGlobalBit));
  return (cls * P + owner) * kTileColours + col;
}
// a solver contact whose colour is final
DBX_D void tile_key_contact(const DevWorld& W, int i, int bA, int bB, int col) {
  int2 br;
  const int bin = tile_classify(W, bA, bB, col, false, &br);
  W.c_tkey[i] = bin; W.c_bref[i] = br; W.c_tcol[i] = -1;
  atomicAdd(&W.t_cur[bin], 1);
}
// every joint (slot order; inactive ones are marked and skipped by the solver)
DBX_D bool joint_active(const DevWorld& W, int j);
DBX_D void tile_key_joints(const DevWorld& W, const int* sjoff) {
  GRID_STRIDE(j, W.nJoints) {
    if (!joint_active(W, j)) { W.j_tkey[j] = -1; W.j_root[j] = -1; continue; }
    const int4 ids = W.j_ids[j];
    // joint slots are colour-major (World::recolourJoints): the colour of slot j is the range of jointColourOff it falls into
    int lo = 0, hi = kMaxJointColours;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (sjoff[mid] <= j) lo = mid; else hi = mid - 1; }
    const bool gear = ids.x == 6;                     // JT_GEAR: the far bodies of joint1 / joint2 are written too (b2gearjoint.d:352-386)
    int2 br;
    const int bin = tile_classify(W, ids.y, ids.z, lo, gear, &br);
    if (gear) {
      const int4 id2 = W.j_ids2[j];
      { const int sl = W.b_tslot[id2.x]; if (sl >= 0) atomicOr(&W.b_xflag[sl], XF_G); }
      { const int sl = W.b_tslot[id2.y]; if (sl >= 0) atomicOr(&W.b_xflag[sl], XF_G); }
    }
    W.j_tkey[j] = bin; W.j_bref[j] = br; W.j_tcol[j] = -1;
    atomicAdd(&W.tj_cur[bin], 1);
  }
}

}  // namespace dbx
