// Minimal stand-in for {fmt}: fmt::format with positional "{}" placeholders (what the samurai demos use for file names and
// the per-iteration log line).  Floating-point values print in shortest round-trip form, like {fmt}.
#pragma once
#include <charconv>
#include <sstream>
#include <string>
#include <type_traits>

namespace fmt
{
    namespace detail
    {
        template <class T>
        inline void put(std::string& out, const T& v)
        {
            if constexpr (std::is_floating_point_v<T>)
            {
                char buf[64];
                auto r = std::to_chars(buf, buf + sizeof(buf), v);
                out.append(buf, r.ptr);
            }
            else
            {
                std::ostringstream os;
                os << v;
                out += os.str();
            }
        }

        inline void format_rec(std::string& out, const char* f)
        {
            out += f;
        }

        template <class T, class... Ts>
        inline void format_rec(std::string& out, const char* f, const T& v, const Ts&... rest)
        {
            for (; *f; ++f)
            {
                if (f[0] == '{' && f[1] == '{')
                {
                    out += '{';
                    ++f;
                }
                else if (f[0] == '{')
                {
                    while (*f && *f != '}')
                    {
                        ++f;
                    }
                    put(out, v);
                    format_rec(out, *f ? f + 1 : f, rest...);
                    return;
                }
                else if (f[0] == '}' && f[1] == '}')
                {
                    out += '}';
                    ++f;
                }
                else
                {
                    out += *f;
                }
            }
        }
    }

    template <class... Ts>
    inline std::string format(const std::string& f, const Ts&... args)
    {
        std::string out;
        detail::format_rec(out, f.c_str(), args...);
        return out;
    }
}
