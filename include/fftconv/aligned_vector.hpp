// include/fftconv/aligned_vector.hpp -- host-side aligned storage.
//
// Same names as the reference (/root/reference/include/fftconv/aligned_vector.hpp:21-74:
// AlignedAllocator<T, Align = 64>, AlignedVector<T, Align = 64>).  On the B200 path
// these only hold HOST staging data; device workspaces live in libfftconv_cuda's
// plan cache.  Independent implementation on std::aligned_alloc semantics via the
// aligned operator new.
#pragma once

#include <cstddef>
#include <limits>
#include <new>
#include <vector>

namespace fftconv {

template <typename T, std::size_t Alignment = 64> class AlignedAllocator {
  static_assert(Alignment >= alignof(T) && (Alignment & (Alignment - 1)) == 0,
                "alignment must be a power of two not smaller than alignof(T)");

public:
  using value_type = T;
  static constexpr std::align_val_t ALIGNMENT{Alignment};
  template <class U> struct rebind {
    using other = AlignedAllocator<U, Alignment>;
  };

  constexpr AlignedAllocator() noexcept = default;
  template <class U> constexpr AlignedAllocator(const AlignedAllocator<U, Alignment> &) noexcept {}

  [[nodiscard]] T *allocate(std::size_t count) {
    if (count > std::numeric_limits<std::size_t>::max() / sizeof(T)) throw std::bad_array_new_length();
    return static_cast<T *>(::operator new[](count * sizeof(T), ALIGNMENT));
  }
  void deallocate(T *ptr, std::size_t) noexcept { ::operator delete[](ptr, ALIGNMENT); }

  template <class U> bool operator==(const AlignedAllocator<U, Alignment> &) const noexcept { return true; }
  template <class U> bool operator!=(const AlignedAllocator<U, Alignment> &) const noexcept { return false; }
};

template <typename T, std::size_t Alignment = 64>
using AlignedVector = std::vector<T, AlignedAllocator<T, Alignment>>;

} // namespace fftconv
