// Minimal stand-in for <xtensor/containers/xfixed.hpp>: the fixed-size small vector the samurai demos use for box corners
// and velocities (xt::xtensor_fixed<double, xt::xshape<dim>>).  Only rank-1 shapes.
#pragma once
#include <array>
#include <cstddef>
#include <initializer_list>
#include <ostream>
#include <type_traits>

namespace xt
{
    template <std::size_t... N>
    struct xshape
    {
    };

    template <class T, class Shape>
    class xtensor_fixed;

    template <class T, std::size_t N>
    class xtensor_fixed<T, xshape<N>>
    {
      public:

        using value_type = T;

        xtensor_fixed()
        {
            m_data.fill(T{});
        }

        xtensor_fixed(std::initializer_list<T> l)
        {
            m_data.fill(T{});
            std::size_t i = 0;
            for (const T& v : l)
            {
                if (i < N)
                {
                    m_data[i++] = v;
                }
            }
        }

        T& operator[](std::size_t i)
        {
            return m_data[i];
        }

        const T& operator[](std::size_t i) const
        {
            return m_data[i];
        }

        T& operator()(std::size_t i)
        {
            return m_data[i];
        }

        const T& operator()(std::size_t i) const
        {
            return m_data[i];
        }

        static constexpr std::size_t size()
        {
            return N;
        }

        void fill(const T& v)
        {
            m_data.fill(v);
        }

        auto begin()
        {
            return m_data.begin();
        }

        auto end()
        {
            return m_data.end();
        }

        auto begin() const
        {
            return m_data.begin();
        }

        auto end() const
        {
            return m_data.end();
        }

        T* data()
        {
            return m_data.data();
        }

        const T* data() const
        {
            return m_data.data();
        }

      private:

        std::array<T, N> m_data;
    };

    // element-wise arithmetic and comparisons, enough for expressions like
    //   xt::all(xt::abs(cell.center() - 0.5) <= 0.5 * length)      (README.md:111-118, tests/test_periodic.cpp:36)
#define XFIXED_BINARY(OP)                                                                                              \
    template <class T, std::size_t N>                                                                                  \
    auto operator OP(const xtensor_fixed<T, xshape<N>>& a, const xtensor_fixed<T, xshape<N>>& b)                       \
    {                                                                                                                  \
        xtensor_fixed<decltype(T{} OP T{}), xshape<N>> r;                                                              \
        for (std::size_t i = 0; i < N; ++i)                                                                            \
        {                                                                                                              \
            r[i] = a[i] OP b[i];                                                                                       \
        }                                                                                                              \
        return r;                                                                                                      \
    }                                                                                                                  \
    template <class T, std::size_t N, class S, class = std::enable_if_t<std::is_arithmetic_v<S>>>                      \
    auto operator OP(const xtensor_fixed<T, xshape<N>>& a, S b)                                                        \
    {                                                                                                                  \
        xtensor_fixed<decltype(T{} OP S{}), xshape<N>> r;                                                              \
        for (std::size_t i = 0; i < N; ++i)                                                                            \
        {                                                                                                              \
            r[i] = a[i] OP b;                                                                                          \
        }                                                                                                              \
        return r;                                                                                                      \
    }                                                                                                                  \
    template <class T, std::size_t N, class S, class = std::enable_if_t<std::is_arithmetic_v<S>>>                      \
    auto operator OP(S a, const xtensor_fixed<T, xshape<N>>& b)                                                        \
    {                                                                                                                  \
        xtensor_fixed<decltype(S{} OP T{}), xshape<N>> r;                                                              \
        for (std::size_t i = 0; i < N; ++i)                                                                            \
        {                                                                                                              \
            r[i] = a OP b[i];                                                                                          \
        }                                                                                                              \
        return r;                                                                                                      \
    }
    XFIXED_BINARY(+)
    XFIXED_BINARY(-)
    XFIXED_BINARY(*)
    XFIXED_BINARY(/)
    XFIXED_BINARY(<=)
    XFIXED_BINARY(<)
    XFIXED_BINARY(>=)
    XFIXED_BINARY(>)
    XFIXED_BINARY(==)
    XFIXED_BINARY(<<)
    XFIXED_BINARY(>>)
#undef XFIXED_BINARY

    template <class T, std::size_t N>
    auto abs(const xtensor_fixed<T, xshape<N>>& a)
    {
        xtensor_fixed<T, xshape<N>> r;
        for (std::size_t i = 0; i < N; ++i)
        {
            r[i] = a[i] < T{} ? -a[i] : a[i];
        }
        return r;
    }

    template <class T, std::size_t N>
    bool all(const xtensor_fixed<T, xshape<N>>& a)
    {
        for (std::size_t i = 0; i < N; ++i)
        {
            if (!a[i])
            {
                return false;
            }
        }
        return true;
    }

    template <class T, std::size_t N>
    bool any(const xtensor_fixed<T, xshape<N>>& a)
    {
        for (std::size_t i = 0; i < N; ++i)
        {
            if (a[i])
            {
                return true;
            }
        }
        return false;
    }

    template <class T, std::size_t N>
    T sum(const xtensor_fixed<T, xshape<N>>& a)
    {
        T s{};
        for (std::size_t i = 0; i < N; ++i)
        {
            s += a[i];
        }
        return s;
    }

    template <class T, std::size_t N>
    std::ostream& operator<<(std::ostream& os, const xtensor_fixed<T, xshape<N>>& v)
    {
        os << "{";
        for (std::size_t i = 0; i < N; ++i)
        {
            os << (i ? ", " : "") << v[i];
        }
        return os << "}";
    }
}
