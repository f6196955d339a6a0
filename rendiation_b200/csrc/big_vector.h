// Vectors that hold a large part of a scene (tree nodes, triangle records, node views): std::vector over an allocator that takes
// blocks of 4 MB and more straight from mmap, 2 MB aligned and marked MADV_HUGEPAGE.  A commit of a million triangles touches half
// a gigabyte of fresh memory; at 4 KB a page the faults alone cost more than building the tree (3 us per fault in the VMs this runs
// in: 65-70 ms per 90 MB, against 17 ms with transparent huge pages — which are given on request only, "madvise" mode).
#pragma once

#include <cstddef>
#include <cstdint>
#include <new>
#include <vector>

#if defined(__linux__)
#include <sys/mman.h>
#endif

namespace rdn {

template <class T>
struct HugeAllocator {
  using value_type = T;
  static constexpr size_t HUGE_MIN = size_t(4) << 20, HUGE_PAGE = size_t(2) << 20;
  HugeAllocator() = default;
  template <class U>
  HugeAllocator(const HugeAllocator<U> &) {}
  static size_t rounded(size_t bytes) { return (bytes + HUGE_PAGE - 1) / HUGE_PAGE * HUGE_PAGE; }
  T *allocate(size_t n) {
    const size_t bytes = n * sizeof(T);
#if defined(__linux__)
    if (bytes >= HUGE_MIN) {
      const size_t len = rounded(bytes);
      void *raw = mmap(nullptr, len + HUGE_PAGE, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
      if (raw == MAP_FAILED) throw std::bad_alloc();
      const uintptr_t p = reinterpret_cast<uintptr_t>(raw), a = (p + HUGE_PAGE - 1) / HUGE_PAGE * HUGE_PAGE;
      if (a > p) munmap(raw, a - p);                                             // (the mapping is trimmed to the aligned block, so that
      if (p + HUGE_PAGE > a) munmap(reinterpret_cast<void *>(a + len), p + HUGE_PAGE - a);  // deallocate needs no header)
      madvise(reinterpret_cast<void *>(a), len, MADV_HUGEPAGE);
      return reinterpret_cast<T *>(a);
    }
#endif
    return static_cast<T *>(::operator new(bytes));
  }
  void deallocate(T *p, size_t n) noexcept {
#if defined(__linux__)
    if (n * sizeof(T) >= HUGE_MIN) { munmap(p, rounded(n * sizeof(T))); return; }
#endif
    ::operator delete(p);
  }
  template <class U>
  bool operator==(const HugeAllocator<U> &) const { return true; }
  template <class U>
  bool operator!=(const HugeAllocator<U> &) const { return false; }
};

template <class T>
using BigVector = std::vector<T, HugeAllocator<T>>;

}  // namespace rdn
