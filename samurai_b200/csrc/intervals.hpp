// Host-side interval containers and eager set algebra for one mesh level.
//
// Role in the reference: include/samurai/interval.hpp:50-64 (Interval), level_cell_array.hpp:249-255
// (LevelCellArray: per-dim interval vectors + offsets) and the lazy set engine under subset/** .
// This is NOT a translation of those: a level is a flat CSR  row-key -> sorted disjoint x-intervals,
// and set expressions are evaluated eagerly into new LevelSets by streaming row merges.  The same CSR
// (row key, x_start, x_end, storage offset) is what gets uploaded to the device.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <limits>
#include <cassert>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace smr
{
    constexpr int KEY_BIAS = 1 << 24;

    inline int64_t mk_key(int y, int z)
    {
        return (static_cast<int64_t>(z + KEY_BIAS) << 32) | static_cast<uint32_t>(y + KEY_BIAS);
    }

    inline int key_y(int64_t k)
    {
        return static_cast<int>(static_cast<uint32_t>(k & 0xffffffffLL)) - KEY_BIAS;
    }

    inline int key_z(int64_t k)
    {
        return static_cast<int>(k >> 32) - KEY_BIAS;
    }

    // One level's cell set.  Rows sorted by (z, y); intervals in a row sorted, disjoint and non-adjacent.
    struct LevelSet
    {
        std::vector<int64_t> key; // row key (y, z)
        std::vector<int32_t> ptr; // rows + 1 entries into xs/xe/off
        std::vector<int32_t> xs;  // interval start
        std::vector<int32_t> xe;  // interval end (exclusive)
        std::vector<int64_t> off; // storage offset of cell xs (filled by assign_offsets / locate)

        LevelSet()
        {
            ptr.push_back(0);
        }

        size_t rows() const
        {
            return key.size();
        }

        size_t n_intervals() const
        {
            return xs.size();
        }

        bool empty() const
        {
            return xs.empty();
        }

        int64_t n_cells() const
        {
            int64_t n = 0;
            for (size_t i = 0; i < xs.size(); ++i)
            {
                n += xe[i] - xs[i];
            }
            return n;
        }

        void clear()
        {
            key.clear();
            ptr.assign(1, 0);
            xs.clear();
            xe.clear();
            off.clear();
        }

        // append an interval to the current last row (must be called with increasing keys / starts)
        void push(int64_t k, int s, int e)
        {
            if (s >= e)
            {
                return;
            }
            if (key.empty() || key.back() != k)
            {
                assert(key.empty() || key.back() < k);
                key.push_back(k);
                ptr.push_back(ptr.back());
            }
            else if (xe.back() >= s && ptr[ptr.size() - 2] < static_cast<int32_t>(xs.size()))
            {
                // overlapping or adjacent with the previous interval of the same row: merge
                xe.back() = std::max(xe.back(), e);
                return;
            }
            xs.push_back(s);
            xe.push_back(e);
            ptr.back() = static_cast<int32_t>(xs.size());
        }

        bool same_cells(const LevelSet& o) const
        {
            return key == o.key && ptr == o.ptr && xs == o.xs && xe == o.xe;
        }

        // row index of key k or -1
        int find_row(int64_t k) const
        {
            auto it = std::lower_bound(key.begin(), key.end(), k);
            if (it == key.end() || *it != k)
            {
                return -1;
            }
            return static_cast<int>(it - key.begin());
        }

        // interval index containing x in row r, or -1
        int find_ivl(int r, int x) const
        {
            int lo = ptr[r], hi = ptr[r + 1];
            // last interval with xs <= x
            int a = lo, b = hi;
            while (a < b)
            {
                int m = (a + b) >> 1;
                if (xs[m] <= x)
                {
                    a = m + 1;
                }
                else
                {
                    b = m;
                }
            }
            int i = a - 1;
            if (i >= lo && x < xe[i])
            {
                return i;
            }
            return -1;
        }

        int find_ivl(int64_t k, int x) const
        {
            int r = find_row(k);
            return r < 0 ? -1 : find_ivl(r, x);
        }

        bool contains(int64_t k, int x) const
        {
            return find_ivl(k, x) >= 0;
        }

        // storage offset of cell (x, row k); requires [x, x_last] inside one interval. -1 if absent.
        int64_t offset_of(int64_t k, int x, int x_last) const
        {
            int i = find_ivl(k, x);
            if (i < 0 || x_last >= xe[i])
            {
                return -1;
            }
            return off[i] + (x - xs[i]);
        }
    };

    // Moving lookup into a LevelSet for queries that arrive in (mostly) non-decreasing (row key, x) order: the row is
    // found by galloping from the previous hit, the interval by advancing a cursor.  Amortised O(1) per query where a
    // fresh find_row/find_ivl pair costs two binary searches.
    struct Probe
    {
        const LevelSet* s = nullptr;
        size_t hint       = 0;
        int row           = -1;
        int q = 0, qe = 0;

        Probe() = default;

        explicit Probe(const LevelSet& ls)
            : s(&ls)
        {
        }

        bool seek(int64_t k)
        {
            const size_t n = s->key.size();
            size_t lo, hi;
            if (hint < n && s->key[hint] <= k)
            {
                if (s->key[hint] == k)
                {
                    lo = hi = hint;
                }
                else
                {
                    size_t step = 1;
                    lo          = hint;
                    while (lo + step < n && s->key[lo + step] < k)
                    {
                        lo += step;
                        step <<= 1;
                    }
                    hi = std::min(n, lo + step + 1);
                    lo = static_cast<size_t>(std::lower_bound(s->key.begin() + static_cast<std::ptrdiff_t>(lo), s->key.begin() + static_cast<std::ptrdiff_t>(hi), k)
                                             - s->key.begin());
                }
            }
            else
            {
                lo = static_cast<size_t>(std::lower_bound(s->key.begin(), s->key.end(), k) - s->key.begin());
            }
            if (lo < n && s->key[lo] == k)
            {
                hint = lo;
                row  = static_cast<int>(lo);
                q    = s->ptr[lo];
                qe   = s->ptr[lo + 1];
                return true;
            }
            hint = std::min(lo, n ? n - 1 : 0);
            row  = -1;
            return false;
        }

        // storage offset of x when [x, x_last] lies inside one interval of the current row, else -1
        int64_t offset(int x, int x_last)
        {
            if (row < 0)
            {
                return -1;
            }
            if (q > s->ptr[row] && s->xs[q > qe - 1 ? qe - 1 : q] > x)
            {
                q = s->ptr[row]; // query went backwards: restart the row
            }
            while (q < qe && s->xe[q] <= x)
            {
                ++q;
            }
            if (q >= qe || s->xs[q] > x || x_last >= s->xe[q])
            {
                return -1;
            }
            return s->off[q] + (x - s->xs[q]);
        }

        bool contains(int x)
        {
            return offset(x, x) >= 0;
        }
    };

    // ------------------------------------------------------------------------------------------------
    // row-level merges on sorted disjoint interval lists
    // ------------------------------------------------------------------------------------------------
    enum class SetOp
    {
        Union,
        Inter,
        Diff
    };

    namespace detail
    {
        inline void row_union(const LevelSet& a, int ra, const LevelSet& b, int rb, int64_t k, LevelSet& out)
        {
            int i = a.ptr[ra], ie = a.ptr[ra + 1], j = b.ptr[rb], je = b.ptr[rb + 1];
            while (i < ie || j < je)
            {
                if (j >= je || (i < ie && a.xs[i] <= b.xs[j]))
                {
                    out.push(k, a.xs[i], a.xe[i]);
                    ++i;
                }
                else
                {
                    out.push(k, b.xs[j], b.xe[j]);
                    ++j;
                }
            }
        }

        inline void row_inter(const LevelSet& a, int ra, const LevelSet& b, int rb, int64_t k, LevelSet& out)
        {
            int i = a.ptr[ra], ie = a.ptr[ra + 1], j = b.ptr[rb], je = b.ptr[rb + 1];
            while (i < ie && j < je)
            {
                int s = std::max(a.xs[i], b.xs[j]);
                int e = std::min(a.xe[i], b.xe[j]);
                if (s < e)
                {
                    out.push(k, s, e);
                }
                if (a.xe[i] < b.xe[j])
                {
                    ++i;
                }
                else
                {
                    ++j;
                }
            }
        }

        inline void row_diff(const LevelSet& a, int ra, const LevelSet& b, int rb, int64_t k, LevelSet& out)
        {
            int i = a.ptr[ra], ie = a.ptr[ra + 1], j = b.ptr[rb], je = b.ptr[rb + 1];
            for (; i < ie; ++i)
            {
                int s = a.xs[i];
                const int e = a.xe[i];
                while (j < je && b.xe[j] <= s)
                {
                    ++j;
                }
                int jj = j;
                while (s < e)
                {
                    if (jj >= je || b.xs[jj] >= e)
                    {
                        out.push(k, s, e);
                        break;
                    }
                    if (b.xs[jj] > s)
                    {
                        out.push(k, s, b.xs[jj]);
                    }
                    s = std::max(s, b.xe[jj]);
                    ++jj;
                }
            }
        }

        inline void copy_row(const LevelSet& a, int ra, int64_t k, LevelSet& out)
        {
            for (int i = a.ptr[ra]; i < a.ptr[ra + 1]; ++i)
            {
                out.push(k, a.xs[i], a.xe[i]);
            }
        }
    }

    template <SetOp OP>
    inline LevelSet set_op(const LevelSet& a, const LevelSet& b)
    {
        LevelSet out;
        if (OP == SetOp::Inter && (a.empty() || b.empty()))
        {
            return out;
        }
        size_t i = 0, j = 0;
        size_t na = a.rows(), nb = b.rows();
        if (OP == SetOp::Inter)
        {
            // only the common key range matters: skip straight to it (chunked callers intersect a small set with a big one)
            i  = static_cast<size_t>(std::lower_bound(a.key.begin(), a.key.end(), b.key.front()) - a.key.begin());
            j  = static_cast<size_t>(std::lower_bound(b.key.begin(), b.key.end(), a.key.front()) - b.key.begin());
            na = static_cast<size_t>(std::upper_bound(a.key.begin(), a.key.end(), b.key.back()) - a.key.begin());
            nb = static_cast<size_t>(std::upper_bound(b.key.begin(), b.key.end(), a.key.back()) - b.key.begin());
            const size_t guess = std::min(na - std::min(i, na), nb - std::min(j, nb));
            out.key.reserve(guess);
            out.xs.reserve(2 * guess);
            out.xe.reserve(2 * guess);
        }
        else
        {
            out.key.reserve(OP == SetOp::Union ? na + nb : na);
            out.xs.reserve(a.xs.size() + (OP == SetOp::Union ? b.xs.size() : 0));
            out.xe.reserve(a.xs.size() + (OP == SetOp::Union ? b.xs.size() : 0));
        }
        while (OP == SetOp::Inter ? (i < na && j < nb) : (i < na || j < nb))
        {
            if (j >= nb || (i < na && a.key[i] < b.key[j]))
            {
                if (OP != SetOp::Inter)
                {
                    detail::copy_row(a, static_cast<int>(i), a.key[i], out);
                }
                ++i;
            }
            else if (i >= na || b.key[j] < a.key[i])
            {
                if (OP == SetOp::Union)
                {
                    detail::copy_row(b, static_cast<int>(j), b.key[j], out);
                }
                ++j;
            }
            else
            {
                const int64_t k = a.key[i];
                if (OP == SetOp::Union)
                {
                    detail::row_union(a, static_cast<int>(i), b, static_cast<int>(j), k, out);
                }
                else if (OP == SetOp::Inter)
                {
                    detail::row_inter(a, static_cast<int>(i), b, static_cast<int>(j), k, out);
                }
                else
                {
                    detail::row_diff(a, static_cast<int>(i), b, static_cast<int>(j), k, out);
                }
                ++i;
                ++j;
            }
        }
        return out;
    }

    inline LevelSet set_union(const LevelSet& a, const LevelSet& b)
    {
        if (a.empty())
        {
            return b;
        }
        if (b.empty())
        {
            return a;
        }
        return set_op<SetOp::Union>(a, b);
    }

    inline LevelSet set_inter(const LevelSet& a, const LevelSet& b)
    {
        return set_op<SetOp::Inter>(a, b);
    }

    inline LevelSet set_diff(const LevelSet& a, const LevelSet& b)
    {
        if (a.empty() || b.empty())
        {
            return a;
        }
        return set_op<SetOp::Diff>(a, b);
    }

    // ------------------------------------------------------------------------------------------------
    // regrouping builder: row fragments (new key, source row, x transform) -> LevelSet
    // ------------------------------------------------------------------------------------------------
    struct Frag
    {
        int64_t key;
        int32_t row;
        int32_t tag; // free for the transform (e.g. which copy)
    };

    // xform(tag, s, e, &ns, &ne): transformed interval
    template <class F>
    inline LevelSet regroup(const LevelSet& src, std::vector<Frag>& frags, bool sorted, F&& xform)
    {
        LevelSet out;
        if (!sorted)
        {
            std::stable_sort(frags.begin(),
                             frags.end(),
                             [](const Frag& a, const Frag& b)
                             {
                                 return a.key < b.key;
                             });
        }
        std::vector<std::pair<int32_t, int32_t>> tmp;
        size_t i = 0;
        while (i < frags.size())
        {
            size_t j = i + 1;
            while (j < frags.size() && frags[j].key == frags[i].key)
            {
                ++j;
            }
            if (j == i + 1)
            {
                const int r = frags[i].row;
                for (int q = src.ptr[r]; q < src.ptr[r + 1]; ++q)
                {
                    int ns, ne;
                    xform(frags[i].tag, src.xs[q], src.xe[q], ns, ne);
                    out.push(frags[i].key, ns, ne);
                }
            }
            else
            {
                tmp.clear();
                for (size_t f = i; f < j; ++f)
                {
                    const int r = frags[f].row;
                    for (int q = src.ptr[r]; q < src.ptr[r + 1]; ++q)
                    {
                        int ns, ne;
                        xform(frags[f].tag, src.xs[q], src.xe[q], ns, ne);
                        tmp.emplace_back(ns, ne);
                    }
                }
                std::sort(tmp.begin(), tmp.end());
                for (auto& p : tmp)
                {
                    out.push(frags[i].key, p.first, p.second);
                }
            }
            i = j;
        }
        return out;
    }

    inline LevelSet translate(const LevelSet& a, int dx, int dy, int dz)
    {
        LevelSet out = a;
        out.off.clear();
        if (dy != 0 || dz != 0)
        {
            for (auto& k : out.key)
            {
                k = mk_key(key_y(k) + dy, key_z(k) + dz);
            }
        }
        if (dx != 0)
        {
            for (auto& v : out.xs)
            {
                v += dx;
            }
            for (auto& v : out.xe)
            {
                v += dx;
            }
        }
        return out;
    }

    // `.on(level - shift)` : interval >> shift (reference interval.hpp:227-233), rows y>>shift, z>>shift
    inline LevelSet coarsen(const LevelSet& a, int shift, int dim)
    {
        if (shift == 0 || a.empty())
        {
            LevelSet o = a;
            o.off.clear();
            return o;
        }
        std::vector<Frag> frags(a.rows());
        bool sorted = true;
        for (size_t r = 0; r < a.rows(); ++r)
        {
            const int y = dim > 1 ? (key_y(a.key[r]) >> shift) : 0;
            const int z = dim > 2 ? (key_z(a.key[r]) >> shift) : 0;
            frags[r] = {mk_key(y, z), static_cast<int32_t>(r), 0};
            if (r > 0 && frags[r].key < frags[r - 1].key)
            {
                sorted = false;
            }
        }
        return regroup(a,
                       frags,
                       sorted,
                       [shift](int, int s, int e, int& ns, int& ne)
                       {
                           ns = s >> shift;
                           ne = ((e - 1) >> shift) + 1;
                       });
    }

    // `.on(level + shift)` : every cell -> all its descendants
    inline LevelSet refine(const LevelSet& a, int shift, int dim)
    {
        if (shift == 0 || a.empty())
        {
            LevelSet o = a;
            o.off.clear();
            return o;
        }
        const int n  = 1 << shift;
        const int ny = dim > 1 ? n : 1;
        const int nz = dim > 2 ? n : 1;
        std::vector<Frag> frags;
        frags.reserve(a.rows() * static_cast<size_t>(ny * nz));
        for (size_t r = 0; r < a.rows(); ++r)
        {
            const int y = key_y(a.key[r]), z = key_z(a.key[r]);
            for (int cz = 0; cz < nz; ++cz)
            {
                for (int cy = 0; cy < ny; ++cy)
                {
                    frags.push_back({mk_key(dim > 1 ? (y << shift) + cy : 0, dim > 2 ? (z << shift) + cz : 0), static_cast<int32_t>(r), 0});
                }
            }
        }
        return regroup(a,
                       frags,
                       dim < 3,
                       [shift](int, int s, int e, int& ns, int& ne)
                       {
                           ns = s << shift;
                           ne = e << shift;
                       });
    }

    // box expansion by w cells in every dimension (reference subset/expansion.hpp, nestedExpand with use_native_expand)
    inline LevelSet expand(const LevelSet& a, int w, int dim)
    {
        if (w == 0 || a.empty())
        {
            LevelSet o = a;
            o.off.clear();
            return o;
        }
        const int wy = dim > 1 ? w : 0;
        const int wz = dim > 2 ? w : 0;
        std::vector<Frag> frags;
        frags.reserve(a.rows() * static_cast<size_t>((2 * wy + 1) * (2 * wz + 1)));
        for (size_t r = 0; r < a.rows(); ++r)
        {
            const int y = key_y(a.key[r]), z = key_z(a.key[r]);
            for (int dz = -wz; dz <= wz; ++dz)
            {
                for (int dy = -wy; dy <= wy; ++dy)
                {
                    frags.push_back({mk_key(y + dy, z + dz), static_cast<int32_t>(r), 0});
                }
            }
        }
        return regroup(a,
                       frags,
                       false,
                       [w](int, int s, int e, int& ns, int& ne)
                       {
                           ns = s - w;
                           ne = e + w;
                       });
    }

    // axis-aligned box [lo, hi) as a LevelSet
    inline LevelSet make_box(int dim, const int lo[3], const int hi[3])
    {
        LevelSet out;
        const int z0 = dim > 2 ? lo[2] : 0, z1 = dim > 2 ? hi[2] : 1;
        const int y0 = dim > 1 ? lo[1] : 0, y1 = dim > 1 ? hi[1] : 1;
        if (lo[0] >= hi[0])
        {
            return out;
        }
        for (int z = z0; z < z1; ++z)
        {
            for (int y = y0; y < y1; ++y)
            {
                out.push(mk_key(y, z), lo[0], hi[0]);
            }
        }
        return out;
    }

    // a ∩ box[lo, hi)   (streaming; the domain pyramid of the reference is always a box here)
    inline LevelSet clip_box(const LevelSet& a, int dim, const int lo[3], const int hi[3])
    {
        LevelSet out;
        for (size_t r = 0; r < a.rows(); ++r)
        {
            const int y = key_y(a.key[r]), z = key_z(a.key[r]);
            if (dim > 1 && (y < lo[1] || y >= hi[1]))
            {
                continue;
            }
            if (dim > 2 && (z < lo[2] || z >= hi[2]))
            {
                continue;
            }
            for (int q = a.ptr[r]; q < a.ptr[r + 1]; ++q)
            {
                out.push(a.key[r], std::max(a.xs[q], lo[0]), std::min(a.xe[q], hi[0]));
            }
        }
        return out;
    }

    // a \ box[lo, hi)
    inline LevelSet minus_box(const LevelSet& a, int dim, const int lo[3], const int hi[3])
    {
        LevelSet out;
        for (size_t r = 0; r < a.rows(); ++r)
        {
            const int y         = key_y(a.key[r]), z = key_z(a.key[r]);
            const bool row_in   = (dim < 2 || (y >= lo[1] && y < hi[1])) && (dim < 3 || (z >= lo[2] && z < hi[2]));
            for (int q = a.ptr[r]; q < a.ptr[r + 1]; ++q)
            {
                if (!row_in)
                {
                    out.push(a.key[r], a.xs[q], a.xe[q]);
                }
                else
                {
                    out.push(a.key[r], a.xs[q], std::min(a.xe[q], lo[0]));
                    out.push(a.key[r], std::max(a.xs[q], hi[0]), a.xe[q]);
                }
            }
        }
        return out;
    }

    // Collects arbitrary (key, s, e) triples and builds a LevelSet (sort + merge).
    struct SetBuilder
    {
        struct T
        {
            int64_t k;
            int32_t s, e;
        };

        std::vector<T> v;

        void add(int64_t k, int s, int e)
        {
            if (s < e)
            {
                v.push_back({k, s, e});
            }
        }

        bool empty() const
        {
            return v.empty();
        }

        LevelSet build()
        {
            std::sort(v.begin(),
                      v.end(),
                      [](const T& a, const T& b)
                      {
                          return a.k < b.k || (a.k == b.k && a.s < b.s);
                      });
            LevelSet out;
            for (auto& t : v)
            {
                out.push(t.k, t.s, t.e);
            }
            return out;
        }
    };

    // ------------------------------------------------------------------------------------------------
    // intra-level parallelism: expand / coarsen / refine / translate / intersection or difference with a fixed set all
    // distribute over the union of their input, so a large level is cut into row chunks that are processed
    // independently and whose results are united again (chunk results overlap only near the cuts)
    // ------------------------------------------------------------------------------------------------
    inline LevelSet slice_rows(const LevelSet& a, size_t r0, size_t r1, bool with_offsets = false)
    {
        LevelSet out;
        if (r0 >= r1)
        {
            return out;
        }
        const int q0 = a.ptr[r0], q1 = a.ptr[r1];
        if (with_offsets && a.off.size() == a.xs.size())
        {
            out.off.assign(a.off.begin() + q0, a.off.begin() + q1);
        }
        out.key.assign(a.key.begin() + static_cast<std::ptrdiff_t>(r0), a.key.begin() + static_cast<std::ptrdiff_t>(r1));
        out.xs.assign(a.xs.begin() + q0, a.xs.begin() + q1);
        out.xe.assign(a.xe.begin() + q0, a.xe.begin() + q1);
        out.ptr.resize(r1 - r0 + 1);
        for (size_t r = r0; r <= r1; ++r)
        {
            out.ptr[r - r0] = a.ptr[r] - q0;
        }
        return out;
    }

    // row boundaries of at most `max_chunks` chunks holding about `target` intervals each
    inline std::vector<size_t> chunk_rows(const LevelSet& a, size_t target, size_t max_chunks)
    {
        // tuning knobs (environment, read once): scale the chunk size / the chunk count limit of every caller
        // (measured on the 16-core box: half the callers' nominal chunk size and twice their chunk count is ~3 % faster per step
        // than the nominal values, a quarter / four times is slower)
        static const double target_scale = std::getenv("SMR_CHUNK_SCALE") ? std::atof(std::getenv("SMR_CHUNK_SCALE")) : 0.5;
        static const long max_override   = std::getenv("SMR_CHUNK_MAX") ? std::atol(std::getenv("SMR_CHUNK_MAX")) : 0;
        target     = std::max<size_t>(1, static_cast<size_t>(static_cast<double>(target) * target_scale));
        max_chunks = max_override > 0 ? static_cast<size_t>(max_override) : 2 * max_chunks;
        std::vector<size_t> cut{0};
        const size_t n = a.n_intervals();
        size_t chunks  = std::min(max_chunks, std::max<size_t>(1, n / std::max<size_t>(target, 1)));
        for (size_t c = 1; c < chunks; ++c)
        {
            const int32_t want = static_cast<int32_t>(n * c / chunks);
            const size_t r     = static_cast<size_t>(std::lower_bound(a.ptr.begin(), a.ptr.end(), want) - a.ptr.begin());
            if (r > cut.back() && r < a.rows())
            {
                cut.push_back(r);
            }
        }
        cut.push_back(a.rows());
        return cut;
    }

    // union of any number of sets: k-way merge over the row keys, runs of rows owned by a single part are copied in bulk
    inline LevelSet union_all(const std::vector<const LevelSet*>& in)
    {
        std::vector<const LevelSet*> parts;
        for (const LevelSet* p : in)
        {
            if (p != nullptr && !p->empty())
            {
                parts.push_back(p);
            }
        }
        LevelSet out;
        if (parts.empty())
        {
            return out;
        }
        if (parts.size() == 1)
        {
            out = *parts[0];
            out.off.clear();
            return out;
        }
        size_t rows = 0, ivls = 0;
        for (const LevelSet* p : parts)
        {
            rows += p->rows();
            ivls += p->n_intervals();
        }
        out.key.reserve(rows);
        out.ptr.reserve(rows + 1);
        out.xs.reserve(ivls);
        out.xe.reserve(ivls);
        const size_t np = parts.size();
        std::vector<size_t> cur(np, 0);
        std::vector<std::pair<int32_t, int32_t>> tmp;
        constexpr int64_t INF = std::numeric_limits<int64_t>::max();
        while (true)
        {
            // smallest and second smallest current keys
            int64_t k1 = INF, k2 = INF;
            size_t p1 = np;
            int ties  = 0;
            for (size_t p = 0; p < np; ++p)
            {
                if (cur[p] >= parts[p]->rows())
                {
                    continue;
                }
                const int64_t k = parts[p]->key[cur[p]];
                if (k < k1)
                {
                    k2 = k1;
                    k1 = k;
                    p1 = p;
                    ties = 1;
                }
                else if (k == k1)
                {
                    ++ties;
                }
                else if (k < k2)
                {
                    k2 = k;
                }
            }
            if (p1 == np)
            {
                break;
            }
            if (ties == 1)
            {
                // rows of part p1 with key < k2: bulk copy
                const LevelSet& a = *parts[p1];
                const size_t r0   = cur[p1];
                const size_t r1   = static_cast<size_t>(std::lower_bound(a.key.begin() + static_cast<std::ptrdiff_t>(r0), a.key.end(), k2) - a.key.begin());
                const int q0 = a.ptr[r0], q1 = a.ptr[r1];
                const int32_t shift = static_cast<int32_t>(out.xs.size()) - q0;
                out.key.insert(out.key.end(), a.key.begin() + static_cast<std::ptrdiff_t>(r0), a.key.begin() + static_cast<std::ptrdiff_t>(r1));
                out.xs.insert(out.xs.end(), a.xs.begin() + q0, a.xs.begin() + q1);
                out.xe.insert(out.xe.end(), a.xe.begin() + q0, a.xe.begin() + q1);
                for (size_t r = r0 + 1; r <= r1; ++r)
                {
                    out.ptr.push_back(a.ptr[r] + shift);
                }
                cur[p1] = r1;
                continue;
            }
            // several parts hold row k1: merge their intervals
            tmp.clear();
            for (size_t p = 0; p < np; ++p)
            {
                if (cur[p] < parts[p]->rows() && parts[p]->key[cur[p]] == k1)
                {
                    const LevelSet& a = *parts[p];
                    for (int q = a.ptr[cur[p]]; q < a.ptr[cur[p] + 1]; ++q)
                    {
                        tmp.emplace_back(a.xs[q], a.xe[q]);
                    }
                    ++cur[p];
                }
            }
            std::sort(tmp.begin(), tmp.end());
            for (auto& iv : tmp)
            {
                out.push(k1, iv.first, iv.second);
            }
        }
        return out;
    }

    inline LevelSet union_all(const std::vector<LevelSet>& in)
    {
        std::vector<const LevelSet*> p;
        p.reserve(in.size());
        for (const LevelSet& s : in)
        {
            p.push_back(&s);
        }
        return union_all(p);
    }

    // f(part) over row chunks of `a` in parallel (no-op inside an enclosing parallel region), results united
    template <class F>
    inline LevelSet par_rows(const LevelSet& a, F&& f, size_t target = 3000, size_t max_chunks = 16)
    {
        const std::vector<size_t> cut = chunk_rows(a, target, max_chunks);
        const int n                   = static_cast<int>(cut.size()) - 1;
        if (n <= 1)
        {
            return f(a);
        }
        std::vector<LevelSet> res(static_cast<size_t>(n));
#pragma omp parallel for schedule(dynamic, 1)
        for (int c = 0; c < n; ++c)
        {
            res[static_cast<size_t>(c)] = f(slice_rows(a, cut[static_cast<size_t>(c)], cut[static_cast<size_t>(c) + 1]));
        }
        return union_all(res);
    }

    // a OP b with both operands cut at the same row keys
    template <SetOp OP>
    inline LevelSet par_set_op(const LevelSet& a, const LevelSet& b, size_t target = 3000, size_t max_chunks = 16)
    {
        const LevelSet& big           = a.n_intervals() >= b.n_intervals() ? a : b;
        const std::vector<size_t> cut = chunk_rows(big, target, max_chunks);
        const int n                   = static_cast<int>(cut.size()) - 1;
        if (n <= 1)
        {
            return set_op<OP>(a, b);
        }
        std::vector<LevelSet> res(static_cast<size_t>(n));
        auto row_of = [](const LevelSet& s, int64_t k)
        {
            return static_cast<size_t>(std::lower_bound(s.key.begin(), s.key.end(), k) - s.key.begin());
        };
#pragma omp parallel for schedule(dynamic, 1)
        for (int c = 0; c < n; ++c)
        {
            const bool first = c == 0, last = c == n - 1;
            const int64_t klo = first ? 0 : big.key[cut[static_cast<size_t>(c)]];
            const int64_t khi = last ? 0 : big.key[cut[static_cast<size_t>(c) + 1]];
            const LevelSet pa = slice_rows(a, first ? 0 : row_of(a, klo), last ? a.rows() : row_of(a, khi));
            const LevelSet pb = slice_rows(b, first ? 0 : row_of(b, klo), last ? b.rows() : row_of(b, khi));
            res[static_cast<size_t>(c)] = set_op<OP>(pa, pb);
        }
        return union_all(res); // chunk results have disjoint key ranges: pure concatenation
    }

    inline LevelSet par_union(const LevelSet& a, const LevelSet& b)
    {
        if (a.empty())
        {
            return b;
        }
        if (b.empty())
        {
            return a;
        }
        return par_set_op<SetOp::Union>(a, b);
    }

    inline LevelSet par_inter(const LevelSet& a, const LevelSet& b)
    {
        if (a.empty() || b.empty())
        {
            return LevelSet();
        }
        return par_set_op<SetOp::Inter>(a, b);
    }

    inline LevelSet par_diff(const LevelSet& a, const LevelSet& b)
    {
        if (a.empty() || b.empty())
        {
            return a;
        }
        return par_set_op<SetOp::Diff>(a, b);
    }

    // Fill sub.off from the containing intervals of `ref` (sub must be a subset of ref; both sorted).
    // Mirrors Mesh_base::renumbering (reference mesh.hpp:894-911): every sub-mesh reuses the reference numbering.
    inline void locate(LevelSet& sub, const LevelSet& ref)
    {
        sub.off.assign(sub.xs.size(), -1);
        // start at the first row of `sub` (chunked callers locate a small slice in a big set)
        size_t rr = sub.rows() ? static_cast<size_t>(std::lower_bound(ref.key.begin(), ref.key.end(), sub.key[0]) - ref.key.begin()) : 0;
        for (size_t r = 0; r < sub.rows(); ++r)
        {
            while (rr < ref.rows() && ref.key[rr] < sub.key[r])
            {
                ++rr;
            }
            if (rr >= ref.rows() || ref.key[rr] != sub.key[r])
            {
                throw std::out_of_range("locate: row not found in reference mesh");
            }
            int q = ref.ptr[rr];
            const int qe = ref.ptr[rr + 1];
            for (int i = sub.ptr[r]; i < sub.ptr[r + 1]; ++i)
            {
                while (q < qe && ref.xe[q] <= sub.xs[i])
                {
                    ++q;
                }
                if (q >= qe || ref.xs[q] > sub.xs[i] || ref.xe[q] < sub.xe[i])
                {
                    throw std::out_of_range("locate: interval not found in reference mesh");
                }
                sub.off[i] = ref.off[q] + (sub.xs[i] - ref.xs[q]);
            }
        }
    }
} // namespace smr
