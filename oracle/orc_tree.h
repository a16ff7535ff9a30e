// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h header).  PARITY UNPINNED.
//
// Incremental dynamic AABB tree + move/pair buffers; restates
//   src/dbox/collision/b2dynamictree.d  (CreateProxy :110-124, DestroyProxy :127-134, MoveProxy :140-184,
//                                        Query :203-237, AllocateNode :516-553, FreeNode :556-564,
//                                        InsertLeaf :566-708, RemoveLeaf :710-769, Balance :773-915)
//   src/dbox/collision/b2broadphase.d   (UpdatePairs :139-195, QueryCallback :271-294, b2PairLessThan :312-325)
// The tree is kept (not replaced by an LBVH) because node ids from its LIFO free list define proxy-id
// order, which decides fixture A/B of every contact.
#pragma once
#include <algorithm>
#include <vector>
#include "orc_math.h"

namespace orc {

constexpr int kNullNode = -1;

struct TreeNode {
  AABB aabb;
  void* userData = nullptr;
  int parentOrNext = kNullNode;
  int child1 = kNullNode, child2 = kNullNode;
  int height = -1;
  bool isLeaf() const { return child1 == kNullNode; }
};

class DynamicTree {
 public:
  DynamicTree() {
    nodes_.resize(16);
    for (int i = 0; i < 15; ++i) { nodes_[i].parentOrNext = i + 1; nodes_[i].height = -1; }
    nodes_[15].parentOrNext = kNullNode; nodes_[15].height = -1;
    freeList_ = 0;
  }
  int createProxy(const AABB& aabb, void* userData) {
    int id = allocateNode();
    V2 r(kAabbExtension, kAabbExtension);
    nodes_[id].aabb.lo = aabb.lo - r;
    nodes_[id].aabb.hi = aabb.hi + r;
    nodes_[id].userData = userData;
    nodes_[id].height = 0;
    insertLeaf(id);
    return id;
  }
  void destroyProxy(int id) { removeLeaf(id); freeNode(id); }
  bool moveProxy(int id, const AABB& aabb, V2 displacement) {
    if (nodes_[id].aabb.contains(aabb)) return false;
    removeLeaf(id);
    AABB b = aabb;
    V2 r(kAabbExtension, kAabbExtension);
    b.lo = b.lo - r;
    b.hi = b.hi + r;
    V2 d = kAabbMultiplier * displacement;
    if (d.x < 0.0f) b.lo.x += d.x; else b.hi.x += d.x;
    if (d.y < 0.0f) b.lo.y += d.y; else b.hi.y += d.y;
    nodes_[id].aabb = b;
    insertLeaf(id);
    return true;
  }
  void* userData(int id) const { return nodes_[id].userData; }
  const AABB& fatAABB(int id) const { return nodes_[id].aabb; }
  // b2dynamictree.d:503-511: every pool node, in use or not
  void shiftOrigin(V2 newOrigin) { for (auto& n : nodes_) { n.aabb.lo -= newOrigin; n.aabb.hi -= newOrigin; } }
  // state import only (tests, bench transplant): give leaf `id` exactly this fat AABB and re-link it so that every ancestor
  // contains its children again (remove + insert, like MoveProxy :140-184 without the extension / displacement rule)
  void importFatAABB(int id, const AABB& a) { removeLeaf(id); nodes_[id].aabb = a; insertLeaf(id); }

  template <class F> void query(F&& cb, const AABB& aabb) const {
    std::vector<int> stack;
    stack.reserve(256);
    stack.push_back(root_);
    while (!stack.empty()) {
      int nodeId = stack.back();
      stack.pop_back();
      if (nodeId == kNullNode) continue;
      const TreeNode* node = &nodes_[nodeId];
      if (overlap(node->aabb, aabb)) {
        if (node->isLeaf()) { if (!cb(nodeId)) return; }
        else { stack.push_back(node->child1); stack.push_back(node->child2); }
      }
    }
  }
  // b2dynamictree.d:237-331.  cb(p1, p2, maxFraction, nodeId) -> 0: stop, < 0: ignore, else new maxFraction
  template <class F> void rayCast(F&& cb, V2 p1, V2 p2, float maxFraction) const {
    V2 r = p2 - p1;
    r.normalize();
    V2 v = cross(1.0f, r);
    V2 abs_v = absv(v);
    AABB segmentAABB;
    { V2 t = p1 + maxFraction * (p2 - p1); segmentAABB.lo = minv(p1, t); segmentAABB.hi = maxv(p1, t); }
    std::vector<int> stack;
    stack.reserve(256);
    stack.push_back(root_);
    while (!stack.empty()) {
      int nodeId = stack.back();
      stack.pop_back();
      if (nodeId == kNullNode) continue;
      const TreeNode* node = &nodes_[nodeId];
      if (overlap(node->aabb, segmentAABB) == false) continue;
      V2 c = 0.5f * (node->aabb.lo + node->aabb.hi);
      V2 h = 0.5f * (node->aabb.hi - node->aabb.lo);
      float separation = absT(dot(v, p1 - c)) - dot(abs_v, h);
      if (separation > 0.0f) continue;
      if (node->isLeaf()) {
        float value = cb(p1, p2, maxFraction, nodeId);
        if (value == 0.0f) return;
        if (value > 0.0f) {
          maxFraction = value;
          V2 t = p1 + maxFraction * (p2 - p1);
          segmentAABB.lo = minv(p1, t); segmentAABB.hi = maxv(p1, t);
        }
      } else { stack.push_back(node->child1); stack.push_back(node->child2); }
    }
  }
  int height() const { return root_ == kNullNode ? 0 : nodes_[root_].height; }

  // b2dynamictree.d:334-352, 939-1011 (Validate) — structural self-check used by the oracle's own tests.
  bool validate() const {
    if (root_ == kNullNode) return true;
    return validateNode(root_, kNullNode);
  }

 private:
  bool validateNode(int i, int parent) const {
    const TreeNode& n = nodes_[i];
    if (n.parentOrNext != parent) return false;
    if (n.isLeaf()) return n.child2 == kNullNode && n.height == 0;
    const TreeNode& a = nodes_[n.child1]; const TreeNode& b = nodes_[n.child2];
    if (n.height != 1 + maxT(a.height, b.height)) return false;
    AABB c; c.combine(a.aabb, b.aabb);
    if (!(c.lo == n.aabb.lo) || !(c.hi == n.aabb.hi)) return false;
    return validateNode(n.child1, i) && validateNode(n.child2, i);
  }
  int allocateNode() {
    if (freeList_ == kNullNode) {
      int old = (int)nodes_.size();
      nodes_.resize(old * 2);
      for (int i = old; i < old * 2 - 1; ++i) { nodes_[i].parentOrNext = i + 1; nodes_[i].height = -1; }
      nodes_[old * 2 - 1].parentOrNext = kNullNode; nodes_[old * 2 - 1].height = -1;
      freeList_ = old;
    }
    int id = freeList_;
    freeList_ = nodes_[id].parentOrNext;
    nodes_[id].parentOrNext = kNullNode;
    nodes_[id].child1 = kNullNode; nodes_[id].child2 = kNullNode;
    nodes_[id].height = 0;
    nodes_[id].userData = nullptr;
    ++nodeCount_;
    return id;
  }
  void freeNode(int id) {
    nodes_[id].parentOrNext = freeList_;
    nodes_[id].height = -1;
    freeList_ = id;
    --nodeCount_;
  }
  void insertLeaf(int leaf) {
    if (root_ == kNullNode) { root_ = leaf; nodes_[root_].parentOrNext = kNullNode; return; }
    AABB leafAABB = nodes_[leaf].aabb;
    int index = root_;
    while (!nodes_[index].isLeaf()) {
      int child1 = nodes_[index].child1, child2 = nodes_[index].child2;
      float area = nodes_[index].aabb.perimeter();
      AABB combined; combined.combine(nodes_[index].aabb, leafAABB);
      float combinedArea = combined.perimeter();
      float cost = 2.0f * combinedArea;
      float inheritanceCost = 2.0f * (combinedArea - area);
      float cost1, cost2;
      if (nodes_[child1].isLeaf()) {
        AABB a; a.combine(leafAABB, nodes_[child1].aabb);
        cost1 = a.perimeter() + inheritanceCost;
      } else {
        AABB a; a.combine(leafAABB, nodes_[child1].aabb);
        float oldArea = nodes_[child1].aabb.perimeter();
        float newArea = a.perimeter();
        cost1 = (newArea - oldArea) + inheritanceCost;
      }
      if (nodes_[child2].isLeaf()) {
        AABB a; a.combine(leafAABB, nodes_[child2].aabb);
        cost2 = a.perimeter() + inheritanceCost;
      } else {
        AABB a; a.combine(leafAABB, nodes_[child2].aabb);
        float oldArea = nodes_[child2].aabb.perimeter();
        float newArea = a.perimeter();
        cost2 = newArea - oldArea + inheritanceCost;
      }
      if (cost < cost1 && cost < cost2) break;
      index = cost1 < cost2 ? child1 : child2;
    }
    int sibling = index;
    int oldParent = nodes_[sibling].parentOrNext;
    int newParent = allocateNode();
    nodes_[newParent].parentOrNext = oldParent;
    nodes_[newParent].userData = nullptr;
    nodes_[newParent].aabb.combine(leafAABB, nodes_[sibling].aabb);
    nodes_[newParent].height = nodes_[sibling].height + 1;
    if (oldParent != kNullNode) {
      if (nodes_[oldParent].child1 == sibling) nodes_[oldParent].child1 = newParent;
      else nodes_[oldParent].child2 = newParent;
    } else {
      root_ = newParent;
    }
    nodes_[newParent].child1 = sibling;
    nodes_[newParent].child2 = leaf;
    nodes_[sibling].parentOrNext = newParent;
    nodes_[leaf].parentOrNext = newParent;
    index = nodes_[leaf].parentOrNext;
    while (index != kNullNode) {
      index = balance(index);
      int child1 = nodes_[index].child1, child2 = nodes_[index].child2;
      nodes_[index].height = 1 + maxT(nodes_[child1].height, nodes_[child2].height);
      nodes_[index].aabb.combine(nodes_[child1].aabb, nodes_[child2].aabb);
      index = nodes_[index].parentOrNext;
    }
  }
  void removeLeaf(int leaf) {
    if (leaf == root_) { root_ = kNullNode; return; }
    int parent = nodes_[leaf].parentOrNext;
    int grandParent = nodes_[parent].parentOrNext;
    int sibling = nodes_[parent].child1 == leaf ? nodes_[parent].child2 : nodes_[parent].child1;
    if (grandParent != kNullNode) {
      if (nodes_[grandParent].child1 == parent) nodes_[grandParent].child1 = sibling;
      else nodes_[grandParent].child2 = sibling;
      nodes_[sibling].parentOrNext = grandParent;
      freeNode(parent);
      int index = grandParent;
      while (index != kNullNode) {
        index = balance(index);
        int child1 = nodes_[index].child1, child2 = nodes_[index].child2;
        nodes_[index].aabb.combine(nodes_[child1].aabb, nodes_[child2].aabb);
        nodes_[index].height = 1 + maxT(nodes_[child1].height, nodes_[child2].height);
        index = nodes_[index].parentOrNext;
      }
    } else {
      root_ = sibling;
      nodes_[sibling].parentOrNext = kNullNode;
      freeNode(parent);
    }
  }
  // One rotation helper covers both mirror cases of b2dynamictree.d:795-912.
  int rotateUp(int iA, int iUp, int iOther, bool upIsChild2) {
    TreeNode* A = &nodes_[iA];
    TreeNode* U = &nodes_[iUp];
    TreeNode* O = &nodes_[iOther];
    int iX = U->child1, iY = U->child2;
    TreeNode* X = &nodes_[iX];
    TreeNode* Y = &nodes_[iY];
    U->child1 = iA;
    U->parentOrNext = A->parentOrNext;
    A->parentOrNext = iUp;
    if (U->parentOrNext != kNullNode) {
      if (nodes_[U->parentOrNext].child1 == iA) nodes_[U->parentOrNext].child1 = iUp;
      else nodes_[U->parentOrNext].child2 = iUp;
    } else {
      root_ = iUp;
    }
    if (X->height > Y->height) {
      U->child2 = iX;
      (upIsChild2 ? A->child2 : A->child1) = iY;
      Y->parentOrNext = iA;
      A->aabb.combine(O->aabb, Y->aabb);
      U->aabb.combine(A->aabb, X->aabb);
      A->height = 1 + maxT(O->height, Y->height);
      U->height = 1 + maxT(A->height, X->height);
    } else {
      U->child2 = iY;
      (upIsChild2 ? A->child2 : A->child1) = iX;
      X->parentOrNext = iA;
      A->aabb.combine(O->aabb, X->aabb);
      U->aabb.combine(A->aabb, Y->aabb);
      A->height = 1 + maxT(O->height, X->height);
      U->height = 1 + maxT(A->height, Y->height);
    }
    return iUp;
  }
  int balance(int iA) {
    TreeNode* A = &nodes_[iA];
    if (A->isLeaf() || A->height < 2) return iA;
    int iB = A->child1, iC = A->child2;
    int bal = nodes_[iC].height - nodes_[iB].height;
    if (bal > 1) return rotateUp(iA, iC, iB, true);
    if (bal < -1) return rotateUp(iA, iB, iC, false);
    return iA;
  }

  std::vector<TreeNode> nodes_;
  int root_ = kNullNode;
  int nodeCount_ = 0;
  int freeList_ = 0;
};

struct Pair { int a, b; };

class BroadPhase {
 public:
  int createProxy(const AABB& aabb, void* userData) {
    int id = tree_.createProxy(aabb, userData);
    ++proxyCount_;
    moveBuffer_.push_back(id);
    return id;
  }
  void destroyProxy(int id) {
    for (int& m : moveBuffer_) if (m == id) m = kNullNode;
    --proxyCount_;
    tree_.destroyProxy(id);
  }
  void moveProxy(int id, const AABB& aabb, V2 displacement) {
    if (tree_.moveProxy(id, aabb, displacement)) moveBuffer_.push_back(id);
  }
  void touchProxy(int id) { moveBuffer_.push_back(id); }
  const AABB& fatAABB(int id) const { return tree_.fatAABB(id); }
  void* userData(int id) const { return tree_.userData(id); }
  bool testOverlap(int a, int b) const { return overlap(tree_.fatAABB(a), tree_.fatAABB(b)); }
  int proxyCount() const { return proxyCount_; }
  int treeHeight() const { return tree_.height(); }
  DynamicTree& tree() { return tree_; }
  const std::vector<int>& moveBuffer() const { return moveBuffer_; }
  void importMoveBuffer(const std::vector<int>& ids) { moveBuffer_ = ids; }   // state import only
  void importFatAABB(int id, const AABB& a) { tree_.importFatAABB(id, a); }
  const std::vector<Pair>& lastPairs() const { return pairBuffer_; }  // sorted, with duplicates (diagnostics)

  template <class F> void updatePairs(F&& addPair) {
    pairBuffer_.clear();
    for (size_t i = 0; i < moveBuffer_.size(); ++i) {
      int q = moveBuffer_[i];
      if (q == kNullNode) continue;
      const AABB fat = tree_.fatAABB(q);
      tree_.query([&](int proxyId) {
        if (proxyId == q) return true;
        pairBuffer_.push_back(Pair{minT(proxyId, q), maxT(proxyId, q)});
        return true;
      }, fat);
    }
    moveBuffer_.clear();
    std::sort(pairBuffer_.begin(), pairBuffer_.end(), [](const Pair& p, const Pair& q) {
      if (p.a < q.a) return true;
      if (p.a == q.a) return p.b < q.b;
      return false;
    });
    size_t i = 0;
    while (i < pairBuffer_.size()) {
      Pair primary = pairBuffer_[i];
      addPair(tree_.userData(primary.a), tree_.userData(primary.b));
      ++i;
      while (i < pairBuffer_.size()) {
        if (pairBuffer_[i].a != primary.a || pairBuffer_[i].b != primary.b) break;
        ++i;
      }
    }
  }

 private:
  DynamicTree tree_;
  int proxyCount_ = 0;
  std::vector<int> moveBuffer_;
  std::vector<Pair> pairBuffer_;
};

}  // namespace orc
