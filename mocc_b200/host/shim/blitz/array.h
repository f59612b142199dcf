// Minimal stand-in for the subset of Blitz++ that MOCC (youngmit/mocc) uses.
//
// BUILD SHIM. Blitz++ is an external dependency of the
// reference (cmake/FindBlitz.cmake, src/util/blitz_typedefs.hpp:19) that is
// not vendored under /root/reference and not installed in this image. This
// header exists so that the UNMODIFIED reference sources can be compiled here (the
// oracle build oracle/Makefile and the plugin build mocc_b200/host/Makefile). Original, written
// against Blitz's documented semantics, not a copy of Blitz:
//   * Array<T,N>: reference-counted storage, row-major, zero-based
//   * copy construction is SHALLOW (a view); operator=(Array) is a DEEP
//     element-wise copy into existing storage; operator=(scalar) fills
//   * Range(a,b) is INCLUSIVE of b; Range::all(); toEnd
//   * operator()(...) with any mix of int / Range arguments yields a view of
//     rank == number of Range arguments (all-int -> element reference)
//   * eager (non-lazy) arithmetic: a+b, a-b, s*a, a*s, a/s, abs, acos, max, ...
#pragma once

#include <array>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <iomanip>
#include <iostream>
#include <iterator>
#include <limits>
#include <memory>
#include <sstream>
#include <string>
#include <type_traits>
#include <vector>

namespace blitz {

const int toEnd     = std::numeric_limits<int>::max();
const int fromStart = std::numeric_limits<int>::min();

class Range {
public:
    Range() : first_(fromStart), last_(toEnd), all_(true)
    {
    }
    Range(int first, int last) : first_(first), last_(last), all_(false)
    {
    }
    explicit Range(int only) : first_(only), last_(only), all_(false)
    {
    }
    static Range all()
    {
        return Range();
    }
    int first(int lo = 0) const
    {
        return (all_ || first_ == fromStart) ? lo : first_;
    }
    // inclusive last index, given the extent of the dimension being sliced
    int last(int extent) const
    {
        return (all_ || last_ == toEnd) ? extent - 1 : last_;
    }

private:
    int first_, last_;
    bool all_;
};

template <class T, int N> class TinyVector {
public:
    TinyVector()
    {
        for (int i = 0; i < N; i++)
            v_[i] = T();
    }
    TinyVector(T a)
    {
        for (int i = 0; i < N; i++)
            v_[i] = a;
    }
    TinyVector(T a, T b)
    {
        static_assert(N == 2, "rank");
        v_[0] = a;
        v_[1] = b;
    }
    TinyVector(T a, T b, T c)
    {
        static_assert(N == 3, "rank");
        v_[0] = a;
        v_[1] = b;
        v_[2] = c;
    }
    TinyVector(T a, T b, T c, T d)
    {
        static_assert(N == 4, "rank");
        v_[0] = a;
        v_[1] = b;
        v_[2] = c;
        v_[3] = d;
    }
    T &operator[](int i)
    {
        return v_[i];
    }
    const T &operator[](int i) const
    {
        return v_[i];
    }
    T &operator()(int i)
    {
        return v_[i];
    }
    const T &operator()(int i) const
    {
        return v_[i];
    }
    bool operator==(const TinyVector &o) const
    {
        for (int i = 0; i < N; i++)
            if (v_[i] != o.v_[i])
                return false;
        return true;
    }

private:
    T v_[N];
};

inline TinyVector<int, 1> shape(int a)
{
    return TinyVector<int, 1>(a);
}
inline TinyVector<int, 2> shape(int a, int b)
{
    return TinyVector<int, 2>(a, b);
}
inline TinyVector<int, 3> shape(int a, int b, int c)
{
    return TinyVector<int, 3>(a, b, c);
}
inline TinyVector<int, 4> shape(int a, int b, int c, int d)
{
    return TinyVector<int, 4>(a, b, c, d);
}

namespace detail {
template <class A> struct is_range : std::false_type {
};
template <> struct is_range<Range> : std::true_type {
};
template <class... A> struct count_ranges;
template <> struct count_ranges<> {
    static const int value = 0;
};
template <class A0, class... A> struct count_ranges<A0, A...> {
    static const int value =
        (is_range<typename std::decay<A0>::type>::value ? 1 : 0) +
        count_ranges<A...>::value;
};
}

template <class T, int N> class Array;

// Strided row-major iterator (forward) over an Array of any rank
template <class T, int N, class Ref> class ArrayIter {
public:
    typedef std::forward_iterator_tag iterator_category;
    typedef T value_type;
    typedef std::ptrdiff_t difference_type;
    typedef typename std::remove_reference<Ref>::type *pointer;
    typedef Ref reference;

    ArrayIter() : base_(nullptr), done_(true)
    {
    }
    ArrayIter(typename std::remove_reference<Ref>::type *base, const int *ext,
              const std::ptrdiff_t *str, bool end)
        : base_(base), done_(end)
    {
        std::size_t n = 1;
        for (int d = 0; d < N; d++) {
            ext_[d] = ext[d];
            str_[d] = str[d];
            idx_[d] = 0;
            n *= (std::size_t)ext[d];
        }
        if (n == 0)
            done_ = true;
    }
    Ref operator*() const
    {
        std::ptrdiff_t off = 0;
        for (int d = 0; d < N; d++)
            off += idx_[d] * str_[d];
        return base_[off];
    }
    pointer operator->() const
    {
        return &(**this);
    }
    ArrayIter &operator++()
    {
        for (int d = N - 1; d >= 0; d--) {
            if (++idx_[d] < ext_[d])
                return *this;
            idx_[d] = 0;
        }
        done_ = true;
        return *this;
    }
    ArrayIter operator++(int)
    {
        ArrayIter t = *this;
        ++(*this);
        return t;
    }
    bool operator==(const ArrayIter &o) const
    {
        if (done_ || o.done_)
            return done_ == o.done_;
        for (int d = 0; d < N; d++)
            if (idx_[d] != o.idx_[d])
                return false;
        return true;
    }
    bool operator!=(const ArrayIter &o) const
    {
        return !(*this == o);
    }

private:
    typename std::remove_reference<Ref>::type *base_;
    int ext_[N];
    std::ptrdiff_t str_[N];
    int idx_[N];
    bool done_;
};

template <class T, int N> class Array {
public:
    typedef T T_numtype;
    typedef ArrayIter<T, N, T &> iterator;
    typedef ArrayIter<T, N, const T &> const_iterator;

    // ---- construction -------------------------------------------------------
    Array() : data_(nullptr)
    {
        for (int d = 0; d < N; d++) {
            ext_[d] = 0;
            str_[d] = 0;
        }
    }
    explicit Array(int n0)
    {
        int e[4] = {n0, 0, 0, 0};
        alloc(e);
    }
    Array(int n0, int n1)
    {
        int e[4] = {n0, n1, 0, 0};
        alloc(e);
    }
    Array(int n0, int n1, int n2)
    {
        int e[4] = {n0, n1, n2, 0};
        alloc(e);
    }
    Array(int n0, int n1, int n2, int n3)
    {
        int e[4] = {n0, n1, n2, n3};
        alloc(e);
    }
    // Blitz accepts a trailing storage-order argument; MOCC passes a stray
    // floating-point literal there (scattering_matrix.cpp:35) which real Blitz
    // converts and effectively ignores. Accept and ignore it.
    template <class F, class = typename std::enable_if<
                           std::is_floating_point<F>::value>::type>
    Array(int n0, int n1, F)
    {
        static_assert(N == 2, "rank");
        int e[4] = {n0, n1, 0, 0};
        alloc(e);
    }
    Array(const TinyVector<int, N> &shp)
    {
        int e[4] = {0, 0, 0, 0};
        for (int d = 0; d < N; d++)
            e[d] = shp[d];
        alloc(e);
    }
    // shallow: shares storage
    Array(const Array &o) : block_(o.block_), data_(o.data_)
    {
        for (int d = 0; d < N; d++) {
            ext_[d] = o.ext_[d];
            str_[d] = o.str_[d];
        }
    }

    // ---- assignment ---------------------------------------------------------
    // deep, element-wise
    Array &operator=(const Array &o)
    {
        if (this->data_ == nullptr && size() == 0 && o.size() != 0) {
            // Blitz would assert on shape mismatch; an empty LHS adopting the
            // RHS shape is the only forgiving case we allow.
            int e[4] = {0, 0, 0, 0};
            for (int d = 0; d < N; d++)
                e[d] = o.ext_[d];
            alloc(e);
        }
        assert(size() == o.size());
        if (overlaps(o)) {
            Array tmp = o.copy();
            assign_from(tmp);
        } else {
            assign_from(o);
        }
        return *this;
    }
    Array &operator=(T v)
    {
        for (auto it = begin(); it != end(); ++it)
            *it = v;
        return *this;
    }

    // ---- shape --------------------------------------------------------------
    void reference(const Array &o)
    {
        block_ = o.block_;
        data_  = o.data_;
        for (int d = 0; d < N; d++) {
            ext_[d] = o.ext_[d];
            str_[d] = o.str_[d];
        }
    }
    void resize(int n0)
    {
        int e[4] = {n0, 0, 0, 0};
        realloc_if_needed(e);
    }
    void resize(int n0, int n1)
    {
        int e[4] = {n0, n1, 0, 0};
        realloc_if_needed(e);
    }
    void resize(int n0, int n1, int n2)
    {
        int e[4] = {n0, n1, n2, 0};
        realloc_if_needed(e);
    }
    void resize(int n0, int n1, int n2, int n3)
    {
        int e[4] = {n0, n1, n2, n3};
        realloc_if_needed(e);
    }
    void resize(const TinyVector<int, N> &shp)
    {
        int e[4] = {0, 0, 0, 0};
        for (int d = 0; d < N; d++)
            e[d] = shp[d];
        realloc_if_needed(e);
    }
    void free()
    {
        block_.reset();
        data_ = nullptr;
        for (int d = 0; d < N; d++) {
            ext_[d] = 0;
            str_[d] = 0;
        }
    }
    int extent(int d) const
    {
        return ext_[d];
    }
    int rows() const
    {
        return ext_[0];
    }
    int cols() const
    {
        return ext_[1];
    }
    int lbound(int) const
    {
        return 0;
    }
    int ubound(int d) const
    {
        return ext_[d] - 1;
    }
    std::size_t size() const
    {
        std::size_t n = 1;
        for (int d = 0; d < N; d++)
            n *= (std::size_t)ext_[d];
        return n;
    }
    std::size_t numElements() const
    {
        return size();
    }
    TinyVector<int, N> shape() const
    {
        TinyVector<int, N> s;
        for (int d = 0; d < N; d++)
            s[d] = ext_[d];
        return s;
    }
    int dimensions() const
    {
        return N;
    }
    static int rank()
    {
        return N;
    }
    std::ptrdiff_t stride(int d) const
    {
        return str_[d];
    }
    bool isStorageContiguous() const
    {
        std::ptrdiff_t expect = 1;
        for (int d = N - 1; d >= 0; d--) {
            if (ext_[d] != 1 && str_[d] != expect)
                return false;
            expect *= ext_[d];
        }
        return true;
    }
    T *data()
    {
        return data_;
    }
    const T *data() const
    {
        return data_;
    }
    T *dataFirst()
    {
        return data_;
    }
    const T *dataFirst() const
    {
        return data_;
    }
    Array copy() const
    {
        Array r(shape());
        r.assign_from(*this);
        return r;
    }

    // ---- iteration (row-major) ---------------------------------------------
    iterator begin()
    {
        return iterator(data_, ext_, str_, false);
    }
    iterator end()
    {
        return iterator(data_, ext_, str_, true);
    }
    const_iterator begin() const
    {
        return const_iterator(data_, ext_, str_, false);
    }
    const_iterator end() const
    {
        return const_iterator(data_, ext_, str_, true);
    }
    const_iterator cbegin() const
    {
        return begin();
    }
    const_iterator cend() const
    {
        return end();
    }

    // ---- element access / slicing ------------------------------------------
    template <class... A>
    typename std::enable_if<detail::count_ranges<A...>::value == 0, T &>::type
    operator()(A... a)
    {
        static_assert(sizeof...(A) == N, "index count != rank");
        return data_[offset(a...)];
    }
    template <class... A>
    typename std::enable_if<detail::count_ranges<A...>::value == 0,
                            const T &>::type
    operator()(A... a) const
    {
        static_assert(sizeof...(A) == N, "index count != rank");
        return data_[offset(a...)];
    }
    template <class... A>
    typename std::enable_if<(detail::count_ranges<A...>::value > 0),
                            Array<T, detail::count_ranges<A...>::value>>::type
    operator()(A... a) const
    {
        static_assert(sizeof...(A) == N, "index count != rank");
        Array<T, detail::count_ranges<A...>::value> r;
        T *p   = data_;
        int od = 0;
        slice_into(r, p, od, 0, a...);
        r.adopt(block_, p);
        return r;
    }
    T &operator[](int i)
    {
        static_assert(N == 1, "operator[] is rank-1 only");
        return data_[i * str_[0]];
    }
    const T &operator[](int i) const
    {
        static_assert(N == 1, "operator[] is rank-1 only");
        return data_[i * str_[0]];
    }

    // ---- compound arithmetic -----------------------------------------------
#define BLITZ_SHIM_COMPOUND(OP)                                                \
    Array &operator OP(const Array &o)                                         \
    {                                                                          \
        assert(size() == o.size());                                            \
        auto src = o.begin();                                                  \
        for (auto it = begin(); it != end(); ++it, ++src)                      \
            *it OP *src;                                                       \
        return *this;                                                          \
    }                                                                          \
    template <class S>                                                         \
    typename std::enable_if<std::is_arithmetic<S>::value, Array &>::type       \
    operator OP(S s)                                                           \
    {                                                                          \
        for (auto it = begin(); it != end(); ++it)                             \
            *it OP s;                                                          \
        return *this;                                                          \
    }
    BLITZ_SHIM_COMPOUND(+=)
    BLITZ_SHIM_COMPOUND(-=)
    BLITZ_SHIM_COMPOUND(*=)
    BLITZ_SHIM_COMPOUND(/=)
#undef BLITZ_SHIM_COMPOUND

    // ---- internals shared between ranks -------------------------------------
    void adopt(const std::shared_ptr<std::vector<T>> &block, T *p)
    {
        block_ = block;
        data_  = p;
    }
    void set_dim(int d, int ext, std::ptrdiff_t str)
    {
        ext_[d] = ext;
        str_[d] = str;
    }

private:
    std::shared_ptr<std::vector<T>> block_;
    T *data_;
    int ext_[N];
    std::ptrdiff_t str_[N];

    void alloc(const int *e)
    {
        std::size_t n = 1;
        for (int d = 0; d < N; d++) {
            ext_[d] = e[d];
            n *= (std::size_t)e[d];
        }
        std::ptrdiff_t s = 1;
        for (int d = N - 1; d >= 0; d--) {
            str_[d] = s;
            s *= ext_[d];
        }
        block_ = std::make_shared<std::vector<T>>(n);
        data_  = block_->data();
    }
    void realloc_if_needed(const int *e)
    {
        bool same = (data_ != nullptr);
        for (int d = 0; d < N; d++)
            same = same && (ext_[d] == e[d]);
        if (!same)
            alloc(e);
    }
    void assign_from(const Array &o)
    {
        auto src = o.begin();
        for (auto it = begin(); it != end(); ++it, ++src)
            *it = *src;
    }
    bool overlaps(const Array &o) const
    {
        return block_ && o.block_ && block_.get() == o.block_.get() &&
               data_ != o.data_;
    }

    std::ptrdiff_t offset() const
    {
        return 0;
    }
    template <class... A> std::ptrdiff_t offset(A... a) const
    {
        const long long idx[] = {(long long)a...};
        std::ptrdiff_t off    = 0;
        for (int d = 0; d < N; d++) {
            assert(idx[d] >= 0 && idx[d] < ext_[d]);
            off += (std::ptrdiff_t)idx[d] * str_[d];
        }
        return off;
    }

    template <int M>
    void slice_into(Array<T, M> &, T *&, int &, int) const
    {
    }
    template <int M, class... A>
    void slice_into(Array<T, M> &r, T *&p, int &od, int d, int i,
                    A... rest) const
    {
        assert(i >= 0 && i < ext_[d]);
        p += (std::ptrdiff_t)i * str_[d];
        slice_into(r, p, od, d + 1, rest...);
    }
    template <int M, class... A>
    void slice_into(Array<T, M> &r, T *&p, int &od, int d, Range rg,
                    A... rest) const
    {
        int f = rg.first(0);
        int l = rg.last(ext_[d]);
        assert(f >= 0 && l < ext_[d]);
        int n = l - f + 1;
        if (n < 0)
            n = 0;
        p += (std::ptrdiff_t)f * str_[d];
        r.set_dim(od, n, str_[d]);
        od++;
        slice_into(r, p, od, d + 1, rest...);
    }
    // other integral index types (size_t, unsigned, long ...)
    template <int M, class I, class... A>
    typename std::enable_if<std::is_integral<I>::value &&
                            !std::is_same<I, int>::value>::type
    slice_into(Array<T, M> &r, T *&p, int &od, int d, I i, A... rest) const
    {
        slice_into(r, p, od, d, (int)i, rest...);
    }
};

// ---- eager expression helpers ----------------------------------------------
#define BLITZ_SHIM_BINARY(OP)                                                  \
    template <class T, int N>                                                  \
    Array<T, N> operator OP(const Array<T, N> &a, const Array<T, N> &b)        \
    {                                                                          \
        assert(a.size() == b.size());                                          \
        Array<T, N> r(a.shape());                                              \
        auto ia = a.begin();                                                   \
        auto ib = b.begin();                                                   \
        for (auto it = r.begin(); it != r.end(); ++it, ++ia, ++ib)             \
            *it = *ia OP * ib;                                                 \
        return r;                                                              \
    }                                                                          \
    template <class T, int N, class S>                                         \
    typename std::enable_if<std::is_arithmetic<S>::value, Array<T, N>>::type   \
    operator OP(const Array<T, N> &a, S s)                                     \
    {                                                                          \
        Array<T, N> r(a.shape());                                              \
        auto ia = a.begin();                                                   \
        for (auto it = r.begin(); it != r.end(); ++it, ++ia)                   \
            *it = *ia OP s;                                                    \
        return r;                                                              \
    }                                                                          \
    template <class T, int N, class S>                                         \
    typename std::enable_if<std::is_arithmetic<S>::value, Array<T, N>>::type   \
    operator OP(S s, const Array<T, N> &a)                                     \
    {                                                                          \
        Array<T, N> r(a.shape());                                              \
        auto ia = a.begin();                                                   \
        for (auto it = r.begin(); it != r.end(); ++it, ++ia)                   \
            *it = s OP * ia;                                                   \
        return r;                                                              \
    }
BLITZ_SHIM_BINARY(+)
BLITZ_SHIM_BINARY(-)
BLITZ_SHIM_BINARY(*)
BLITZ_SHIM_BINARY(/)
#undef BLITZ_SHIM_BINARY

template <class T, int N> Array<T, N> operator-(const Array<T, N> &a)
{
    Array<T, N> r(a.shape());
    auto ia = a.begin();
    for (auto it = r.begin(); it != r.end(); ++it, ++ia)
        *it = -*ia;
    return r;
}

#define BLITZ_SHIM_UNARY(NAME, EXPR)                                           \
    template <class T, int N> Array<T, N> NAME(const Array<T, N> &a)           \
    {                                                                          \
        Array<T, N> r(a.shape());                                              \
        auto ia = a.begin();                                                   \
        for (auto it = r.begin(); it != r.end(); ++it, ++ia) {                 \
            const T &x = *ia;                                                  \
            *it        = EXPR;                                                 \
        }                                                                      \
        return r;                                                              \
    }
BLITZ_SHIM_UNARY(abs, std::abs(x))
BLITZ_SHIM_UNARY(acos, std::acos(x))
BLITZ_SHIM_UNARY(sqrt, std::sqrt(x))
BLITZ_SHIM_UNARY(exp, std::exp(x))
BLITZ_SHIM_UNARY(log, std::log(x))
BLITZ_SHIM_UNARY(sqr, x *x)
BLITZ_SHIM_UNARY(pow2, x *x)
#undef BLITZ_SHIM_UNARY

template <class T, int N> T max(const Array<T, N> &a)
{
    auto it = a.begin();
    T m     = *it;
    for (; it != a.end(); ++it)
        if (*it > m)
            m = *it;
    return m;
}
template <class T, int N> T min(const Array<T, N> &a)
{
    auto it = a.begin();
    T m     = *it;
    for (; it != a.end(); ++it)
        if (*it < m)
            m = *it;
    return m;
}
template <class T, int N> T sum(const Array<T, N> &a)
{
    T s = T();
    for (auto it = a.begin(); it != a.end(); ++it)
        s += *it;
    return s;
}

template <class T> std::ostream &operator<<(std::ostream &os, const Array<T, 1> &a)
{
    os << "(0," << a.extent(0) - 1 << ")\n[ ";
    for (int i = 0; i < a.extent(0); i++)
        os << a(i) << " ";
    os << "]\n";
    return os;
}
template <class T> std::ostream &operator<<(std::ostream &os, const Array<T, 2> &a)
{
    os << "(0," << a.extent(0) - 1 << ") x (0," << a.extent(1) - 1 << ")\n[ ";
    for (int i = 0; i < a.extent(0); i++) {
        for (int j = 0; j < a.extent(1); j++)
            os << a(i, j) << " ";
        if (i != a.extent(0) - 1)
            os << "\n  ";
    }
    os << "]\n";
    return os;
}
template <class T, int N>
typename std::enable_if<(N > 2), std::ostream &>::type
operator<<(std::ostream &os, const Array<T, N> &a)
{
    os << "[ ";
    for (auto it = a.begin(); it != a.end(); ++it)
        os << *it << " ";
    os << "]\n";
    return os;
}

} // namespace blitz
