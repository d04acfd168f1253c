"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain Python) of the reference's brute-force neighbour search.
Only tests/ may import this; nothing under smartcore_b200/ does.

Follows, line by line:
  HeapSelection                /root/reference/src/algorithm/sort/heap_select.rs:7-95
  LinearKNNSearch::find        /root/reference/src/algorithm/neighbour/linear_search.rs:52-84
  LinearKNNSearch::find_radius /root/reference/src/algorithm/neighbour/linear_search.rs:89-110
  Euclidian::distance          /root/reference/src/metrics/distance/euclidian.rs:51-76 (via oracle_py.squared_distance)

PARITY PIN: the reference's own known-answer tests -- heap_select.rs test_add ([2, 0, -5]), test_add1
([0, -1, -5]), test_add2 ([5.6568, 2.8284, 0.0]), test_add_ordered ([3, 2, 1]) and linear_search.rs knn_find
(indices {0,1,2} around 2 in 1..10; radius 3 around 5 -> 2..8; {1,2,3} around [3,3]) -- are checked in
tests/test_oracle.py.
"""
import math


class HeapSelection:
    def __init__(self, k):
        self.k, self.n, self.sorted, self.heap = k, 0, False, []

    def add(self, element, less=lambda a, b: a < b):
        self.sorted = False
        if self.n < self.k:
            self.heap.append(element)
            self.n += 1
            if self.n == self.k:
                self.sort()
        else:
            self.n += 1
            if element < self.heap[0]:
                self.heap[0] = element
                self.sift_down(0, self.k - 1)

    def heapify(self):
        n = len(self.heap)
        if n <= 1:
            return
        for i in range(n // 2 - 1, -1, -1):
            self.sift_down(i, n - 1)

    def sift_down(self, k, n):
        kk = k
        while 2 * kk <= n:
            j = 2 * kk
            if j < n and self.heap[j] < self.heap[j + 1]:
                j += 1
            if self.heap[kk] >= self.heap[j]:            # Equal or Greater (NaN compares as neither: loop goes on)
                break
            self.heap[kk], self.heap[j] = self.heap[j], self.heap[kk]
            kk = j

    def sort(self):
        self.sorted = True
        self.heap.sort(reverse=True)                     # sort_by(|a, b| b.partial_cmp(a))

    def get(self):
        return self.heap


class _KNNPoint:
    """KNNPoint: ordered and compared by distance only (linear_search.rs:113-131)."""
    __slots__ = ("distance", "index")

    def __init__(self, distance, index):
        self.distance, self.index = distance, index

    def __lt__(self, o): return self.distance < o.distance
    def __ge__(self, o): return self.distance >= o.distance
    def __eq__(self, o): return self.distance == o.distance


def find(data, distance, frm, k):
    """LinearKNNSearch::find -> list of (index, distance) in the reference's own (heap) order."""
    if k < 1 or k > len(data):
        raise ValueError("k should be >= 1 and <= length(data)")
    heap = HeapSelection(k)
    for _ in range(k):
        heap.add(_KNNPoint(math.inf, None))
    for i in range(len(data)):
        d = distance(frm, data[i])
        datum = heap.heap[0]                             # peek_mut
        if d < datum.distance:
            datum.distance = d
            datum.index = i
            heap.heapify()
    return [(p.index, p.distance) for p in heap.get() if p.index is not None]


def find_radius(data, distance, frm, radius):
    if radius <= 0:
        raise ValueError("radius should be > 0")
    return [(i, distance(frm, data[i])) for i in range(len(data)) if distance(frm, data[i]) <= radius]
