// Minimal stand-in for <xtensor/containers/xfixed.hpp>: the fixed-size small vector the samurai demos use for box corners
// and velocities (xt::xtensor_fixed<double, xt::xshape<dim>>).  Only rank-1 shapes.
#pragma once
#include <array>
#include <cstddef>
#include <initializer_list>
#include <ostream>

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
