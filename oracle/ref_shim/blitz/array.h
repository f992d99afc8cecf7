// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for the Blitz++ array
// header so that the UNMODIFIED reference sources under /root/reference/src
// compile into oracle/_ref/ without the un-vendored pkdgrav3 checkout.
//
// Blitz++ supplies storage and slicing only on this path (no arithmetic), so
// this shim implements exactly the calls the reference makes:
//   Array<T,1>/<T,2> construction (owning / pre-existing memory / storage
//   order), shallow copy ("reference") semantics, operator()(i), (i,j),
//   (Range), (Range,int), Range(a,b) INCLUSIVE of b, Range::all(), data(),
//   rows(), columns(), reference(), GeneralArrayStorage<2> with the
//   comma-initialiser syntax, shape(), deleteDataWhenDone.
// Call sites: init.cu:32-45,63,70,76,108,134; makeAxis.cpp:22-26;
// countLeft.cpp:25-26; partition.cpp:10-16; orbit.cpp:79,111;
// copyParticles.cu:30-32; partitionGPU.cu:531-533; pst.h:8-12.
#ifndef ORB_REF_SHIM_BLITZ_ARRAY_H
#define ORB_REF_SHIM_BLITZ_ARRAY_H

// The real Blitz++/mdl2 headers pull in most of the standard library; the
// reference relies on that transitively (e.g. std::chrono in orbit.cpp:87).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <iostream>
#include <string>
#include <tuple>
#include <type_traits>
#include <cassert>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>

namespace blitz {

enum preexistingMemoryPolicy { duplicateData, deleteDataWhenDone, neverDeleteData };

class Range {
public:
    int lo, hi;
    bool everything;
    Range(int lo_, int hi_) : lo(lo_), hi(hi_), everything(false) {}
    static Range all() { Range r(0, -1); r.everything = true; return r; }
};

template <class T, int N>
class TinyVector {
    T v[N];
    class ListInit {
        T *next;
    public:
        explicit ListInit(T *n) : next(n) {}
        ListInit operator,(T x) { *next = x; return ListInit(next + 1); }
    };
public:
    TinyVector() { for (int i = 0; i < N; ++i) v[i] = T(); }
    T &operator()(int i) { return v[i]; }
    const T &operator()(int i) const { return v[i]; }
    T &operator[](int i) { return v[i]; }
    const T &operator[](int i) const { return v[i]; }
    // `tv = a, b;` fills elements in order, like Blitz's list initialiser.
    ListInit operator=(T x) { v[0] = x; return ListInit(v + 1); }
};

inline TinyVector<int, 2> shape(int a, int b) {
    TinyVector<int, 2> s;
    s[0] = a; s[1] = b;
    return s;
}

template <int N>
class GeneralArrayStorage {
    TinyVector<int, N> ordering_;      // ordering_(0) = fastest-varying dimension
    TinyVector<int, N> base_;
    TinyVector<bool, N> ascending_;
public:
    GeneralArrayStorage() {
        // Blitz default is C (row-major) order: last dimension fastest.
        for (int i = 0; i < N; ++i) { ordering_[i] = N - 1 - i; base_[i] = 0; ascending_[i] = true; }
    }
    TinyVector<int, N> &ordering() { return ordering_; }
    const TinyVector<int, N> &ordering() const { return ordering_; }
    TinyVector<int, N> &base() { return base_; }
    TinyVector<bool, N> &ascendingFlag() { return ascending_; }
};

template <class T, int N> class Array;

// ---------------------------------------------------------------- rank 1
template <class T>
class Array<T, 1> {
    T *data_;
    int n_;
    std::ptrdiff_t stride_;
public:
    Array() : data_(nullptr), n_(0), stride_(1) {}
    explicit Array(int n) : data_(nullptr), n_(n), stride_(1) {
        // Storage is never released (process-lifetime buffers): keeps Array copies
        // trivially cheap, like Blitz's non-atomic reference counts, so that the
        // by-value Array in partition.cpp:10 costs what it costs with the real library.
        data_ = static_cast<T *>(std::calloc((n > 0 ? (size_t)n : 1) + 16, sizeof(T)));   // +16: absorbs the reference's one-element overrun in MakeAxis (makeAxis.cpp:23-28)
    }
    // Pre-existing memory. The reference hands over cudaMallocHost memory with
    // deleteDataWhenDone; the shim never frees it (process-lifetime buffers).
    Array(T *data, int n, preexistingMemoryPolicy) : data_(data), n_(n), stride_(1) {}
    // view constructor used by the slicing operators
    Array(T *data, int n, std::ptrdiff_t stride) : data_(data), n_(n), stride_(stride) {}
    // Copy = shallow reference to the same storage (Blitz semantics).
    Array(const Array &) = default;
    Array &operator=(const Array &) = delete;   // Blitz would deep-copy; unused by the reference

    void reference(const Array &o) { data_ = o.data_; n_ = o.n_; stride_ = o.stride_; }

    T &operator()(int i) { return data_[(std::ptrdiff_t)i * stride_]; }
    const T &operator()(int i) const { return data_[(std::ptrdiff_t)i * stride_]; }
    Array operator()(const Range &r) const {
        if (r.everything) return Array(data_, n_, stride_);
        return Array(data_ + (std::ptrdiff_t)r.lo * stride_, r.hi - r.lo + 1, stride_);
    }
    T *data() { return data_; }
    const T *data() const { return data_; }
    int rows() const { return n_; }
    int extent(int) const { return n_; }
    int size() const { return n_; }
};

// ---------------------------------------------------------------- rank 2
template <class T>
class Array<T, 2> {
    T *data_;
    int ext_[2];
    std::ptrdiff_t stride_[2];
    void setStrides(const GeneralArrayStorage<2> &st) {
        int fast = st.ordering()(0);
        int slow = st.ordering()(1);
        stride_[fast] = 1;
        stride_[slow] = ext_[fast];
    }
public:
    Array() : data_(nullptr) { ext_[0] = ext_[1] = 0; stride_[0] = stride_[1] = 0; }
    Array(int rows, int cols) {
        ext_[0] = rows; ext_[1] = cols;
        setStrides(GeneralArrayStorage<2>());
        size_t n = (size_t)rows * (size_t)cols;
        data_ = static_cast<T *>(std::calloc(n ? n : 1, sizeof(T)));
    }
    Array(T *data, const TinyVector<int, 2> &shp, preexistingMemoryPolicy,
          const GeneralArrayStorage<2> &st = GeneralArrayStorage<2>()) : data_(data) {
        ext_[0] = shp[0]; ext_[1] = shp[1];
        setStrides(st);
    }
    Array(const Array &) = default;
    Array &operator=(const Array &) = delete;

    void reference(const Array &o) {
        data_ = o.data_;
        ext_[0] = o.ext_[0]; ext_[1] = o.ext_[1];
        stride_[0] = o.stride_[0]; stride_[1] = o.stride_[1];
    }
    T &operator()(int i, int j) { return data_[i * stride_[0] + j * stride_[1]]; }
    const T &operator()(int i, int j) const { return data_[i * stride_[0] + j * stride_[1]]; }
    // column slice: (Range over rows, fixed column)
    Array<T, 1> operator()(const Range &r, int j) const {
        if (r.everything) return Array<T, 1>(data_ + j * stride_[1], ext_[0], stride_[0]);
        return Array<T, 1>(data_ + r.lo * stride_[0] + j * stride_[1], r.hi - r.lo + 1, stride_[0]);
    }
    T *data() { return data_; }
    const T *data() const { return data_; }
    int rows() const { return ext_[0]; }
    int columns() const { return ext_[1]; }
    int extent(int d) const { return ext_[d]; }
};

}  // namespace blitz

#endif
