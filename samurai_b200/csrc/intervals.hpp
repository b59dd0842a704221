// Host-side interval containers and eager set algebra for one mesh level.
//
// Role in the reference: include/samurai/interval.hpp:50-64 (Interval), level_cell_array.hpp:249-255
// (LevelCellArray: per-dim interval vectors + offsets) and the lazy set engine under subset/** .
// This is NOT a translation of those: a level is a flat CSR  row-key -> sorted disjoint x-intervals,
// and set expressions are evaluated eagerly into new LevelSets by streaming row merges.  The same CSR
// (row key, x_start, x_end, storage offset) is what gets uploaded to the device.
#pragma once
#include <algorithm>
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
        const size_t na = a.rows(), nb = b.rows();
        out.key.reserve(OP == SetOp::Union ? na + nb : na);
        out.xs.reserve(a.xs.size() + (OP == SetOp::Union ? b.xs.size() : 0));
        out.xe.reserve(a.xs.size() + (OP == SetOp::Union ? b.xs.size() : 0));
        while (i < na || j < nb)
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

    // Fill sub.off from the containing intervals of `ref` (sub must be a subset of ref; both sorted).
    // Mirrors Mesh_base::renumbering (reference mesh.hpp:894-911): every sub-mesh reuses the reference numbering.
    inline void locate(LevelSet& sub, const LevelSet& ref)
    {
        sub.off.assign(sub.xs.size(), -1);
        size_t rr = 0;
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
