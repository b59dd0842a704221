// Minimal HDF5 writer / reader for the files samurai's save() produces (io/hdf5.hpp:680-950): no HDF5 library exists in this
// environment, so the on-disk structures are written directly -- superblock version 0, old-style groups (symbol table: B-tree v1
// + SNOD + local heap), version-1 object headers, contiguous layout version 3, IEEE float64 / unsigned and signed 64-bit
// little-endian datasets.  These are exactly the structures the reference's own golden files use (checked byte pattern by byte
// pattern against tests/reference/finite_volume/*.h5), so h5py / h5diff / ParaView read them like the reference's output.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace samurai::b200::h5
{
    constexpr uint64_t UNDEF = ~uint64_t(0);

    enum class Type
    {
        f64,
        u64,
        i64
    };

    class Writer
    {
      public:

        // datasets are registered with their full path ("/mesh/fields/u"); groups are created on the way
        void add(const std::string& path, Type type, const std::vector<uint64_t>& shape, const void* data)
        {
            uint64_t n = 1;
            for (uint64_t s : shape)
            {
                n *= s;
            }
            Node* g = &m_root;
            std::size_t pos = 1;
            while (true)
            {
                const std::size_t slash = path.find('/', pos);
                const std::string name  = path.substr(pos, slash == std::string::npos ? std::string::npos : slash - pos);
                auto& child             = g->children[name];
                if (!child)
                {
                    child = std::make_unique<Node>();
                }
                g = child.get();
                if (slash == std::string::npos)
                {
                    break;
                }
                pos = slash + 1;
            }
            g->is_dataset = true;
            g->type       = type;
            g->shape      = shape;
            g->data.assign(static_cast<const uint8_t*>(data), static_cast<const uint8_t*>(data) + n * 8);
        }

        void add_scalar(const std::string& path, uint64_t v)
        {
            add(path, Type::u64, {}, &v);
        }

        void add_scalar(const std::string& path, double v)
        {
            add(path, Type::f64, {}, &v);
        }

        void write(const std::string& file)
        {
            m_buf.assign(96, 0);
            layout_group(m_root, 96);
            // superblock (version 0)
            static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
            std::memcpy(m_buf.data(), sig, 8);
            m_buf[13] = 8; // size of offsets
            m_buf[14] = 8; // size of lengths
            put16(16, LEAF_K);
            put16(18, INTERNAL_K);
            put64(24, 0);     // base address
            put64(32, UNDEF); // free-space info
            put64(40, m_buf.size());
            put64(48, UNDEF); // driver info
            // root symbol table entry
            put64(56, 0);
            put64(64, m_root.header);
            put32(72, 1); // cache type 1: group, scratch = B-tree + heap addresses
            put64(80, m_root.btree);
            put64(88, m_root.heap);
            std::ofstream os(file, std::ios::binary | std::ios::trunc);
            if (!os)
            {
                throw std::runtime_error("cannot open '" + file + "' for writing");
            }
            os.write(reinterpret_cast<const char*>(m_buf.data()), static_cast<std::streamsize>(m_buf.size()));
        }

      private:

        static constexpr int LEAF_K     = 32; // up to 64 links per group (one symbol node)
        static constexpr int INTERNAL_K = 16;

        struct Node
        {
            std::map<std::string, std::unique_ptr<Node>> children; // sorted by name, as the symbol node must be
            bool is_dataset = false;
            Type type       = Type::f64;
            std::vector<uint64_t> shape;
            std::vector<uint8_t> data;
            uint64_t header = 0, btree = 0, heap = 0;
        };

        uint64_t alloc(std::size_t n)
        {
            const uint64_t at = (m_buf.size() + 7) & ~uint64_t(7);
            m_buf.resize(at + n, 0);
            return at;
        }

        void put16(uint64_t at, uint16_t v)
        {
            std::memcpy(m_buf.data() + at, &v, 2);
        }

        void put32(uint64_t at, uint32_t v)
        {
            std::memcpy(m_buf.data() + at, &v, 4);
        }

        void put64(uint64_t at, uint64_t v)
        {
            std::memcpy(m_buf.data() + at, &v, 8);
        }

        // message header: type(2) size(2) flags(1) reserved(3), body padded to 8 bytes
        uint64_t message(uint64_t at, uint16_t type, uint16_t size, uint8_t flags)
        {
            put16(at, type);
            put16(at + 2, size);
            m_buf[at + 4] = flags;
            return at + 8;
        }

        void layout_group(Node& g, uint64_t header_at)
        {
            if (g.children.size() > 2 * LEAF_K)
            {
                throw std::runtime_error("too many links in one HDF5 group for this writer");
            }
            // object header (version 1): one symbol-table message
            g.header = header_at == 0 ? alloc(40) : header_at;
            if (header_at != 0)
            {
                m_buf.resize(std::max<std::size_t>(m_buf.size(), header_at + 40), 0);
            }
            // local heap: names, 8-byte padded; offset 0 holds the empty string
            std::vector<uint64_t> name_off;
            std::vector<uint8_t> heap_data(8, 0);
            for (auto& kv : g.children)
            {
                name_off.push_back(heap_data.size());
                heap_data.insert(heap_data.end(), kv.first.begin(), kv.first.end());
                heap_data.push_back(0);
                while (heap_data.size() % 8)
                {
                    heap_data.push_back(0);
                }
            }
            g.btree             = alloc(24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8);
            g.heap              = alloc(32);
            const uint64_t hdat = alloc(heap_data.size());
            const uint64_t snod = alloc(8 + 2 * LEAF_K * 40);
            // children first need their own addresses
            std::vector<uint64_t> child_hdr;
            for (auto& kv : g.children)
            {
                Node& c = *kv.second;
                if (c.is_dataset)
                {
                    layout_dataset(c);
                }
                else
                {
                    layout_group(c, 0);
                }
                child_hdr.push_back(c.header);
            }
            // header
            m_buf[g.header] = 1;  // version
            put16(g.header + 2, 1); // messages
            put32(g.header + 4, 1); // reference count
            put32(g.header + 8, 24); // header data size
            uint64_t body = message(g.header + 16, 0x0011, 16, 0);
            put64(body, g.btree);
            put64(body + 8, g.heap);
            // heap
            std::memcpy(m_buf.data() + g.heap, "HEAP", 4);
            put64(g.heap + 8, heap_data.size());
            put64(g.heap + 16, 1); // no free block (H5HL_FREE_NULL)
            put64(g.heap + 24, hdat);
            std::memcpy(m_buf.data() + hdat, heap_data.data(), heap_data.size());
            // B-tree: one leaf entry pointing at the symbol node
            std::memcpy(m_buf.data() + g.btree, "TREE", 4);
            m_buf[g.btree + 4] = 0; // group node
            m_buf[g.btree + 5] = 0; // level
            put16(g.btree + 6, g.children.empty() ? 0 : 1);
            put64(g.btree + 8, UNDEF);
            put64(g.btree + 16, UNDEF);
            put64(g.btree + 24, 0); // key 0: the empty name
            put64(g.btree + 32, snod);
            put64(g.btree + 40, name_off.empty() ? 0 : name_off.back()); // key 1: the largest name in the node
            // symbol node
            std::memcpy(m_buf.data() + snod, "SNOD", 4);
            m_buf[snod + 4] = 1;
            put16(snod + 6, static_cast<uint16_t>(g.children.size()));
            std::size_t k = 0;
            for (auto& kv : g.children)
            {
                const uint64_t e = snod + 8 + 40 * k;
                put64(e, name_off[k]);
                put64(e + 8, child_hdr[k]);
                if (!kv.second->is_dataset)
                {
                    put32(e + 16, 1);
                    put64(e + 24, kv.second->btree);
                    put64(e + 32, kv.second->heap);
                }
                ++k;
            }
        }

        void layout_dataset(Node& d)
        {
            const std::size_t rank  = d.shape.size();
            const uint16_t space_sz = static_cast<uint16_t>(8 + 16 * rank);
            const uint16_t type_sz  = d.type == Type::f64 ? 24 : 16;
            const std::size_t total = 16 + (8 + space_sz) + (8 + type_sz) + (8 + 8) + (8 + 24);
            d.header                = alloc(total);
            const uint64_t raw      = d.data.empty() ? UNDEF : alloc(d.data.size());
            if (!d.data.empty())
            {
                std::memcpy(m_buf.data() + raw, d.data.data(), d.data.size());
            }
            m_buf[d.header] = 1;
            put16(d.header + 2, 4);
            put32(d.header + 4, 1);
            put32(d.header + 8, static_cast<uint32_t>(total - 16));
            // dataspace, version 1, maximum sizes present (= current sizes)
            uint64_t b = message(d.header + 16, 0x0001, space_sz, 0);
            m_buf[b]     = 1;
            m_buf[b + 1] = static_cast<uint8_t>(rank);
            m_buf[b + 2] = 1;
            for (std::size_t i = 0; i < rank; ++i)
            {
                put64(b + 8 + 8 * i, d.shape[i]);
                put64(b + 8 + 8 * rank + 8 * i, d.shape[i]);
            }
            // datatype (constant message)
            b = message(b + space_sz, 0x0003, type_sz, 1);
            if (d.type == Type::f64)
            {
                static const uint8_t f64[24] = {0x11, 0x20, 0x3f, 0x00, 8, 0, 0, 0, 0, 0, 0x40, 0, 0x34, 0x0b, 0, 0x34, 0xff, 0x03, 0, 0, 0, 0, 0, 0};
                std::memcpy(m_buf.data() + b, f64, 24);
            }
            else
            {
                const uint8_t i64[16] = {0x10, static_cast<uint8_t>(d.type == Type::i64 ? 0x08 : 0x00), 0, 0, 8, 0, 0, 0, 0, 0, 0x40, 0, 0, 0, 0, 0};
                std::memcpy(m_buf.data() + b, i64, 16);
            }
            // fill value: version 2, allocation late, write if set, undefined
            b = message(b + type_sz, 0x0005, 8, 1);
            static const uint8_t fill[8] = {2, 2, 2, 1, 0, 0, 0, 0};
            std::memcpy(m_buf.data() + b, fill, 8);
            // layout: version 3, contiguous
            b            = message(b + 8, 0x0008, 24, 0);
            m_buf[b]     = 3;
            m_buf[b + 1] = 1;
            put64(b + 2, raw);
            put64(b + 10, d.data.size());
        }

        Node m_root;
        std::vector<uint8_t> m_buf;
    };

    // Reader for the same subset (used by samurai::load on files written by dump()).
    class Reader
    {
      public:

        explicit Reader(const std::string& file)
        {
            std::ifstream is(file, std::ios::binary);
            if (!is)
            {
                throw std::runtime_error("cannot open '" + file + "'");
            }
            m_buf.assign(std::istreambuf_iterator<char>(is), std::istreambuf_iterator<char>());
            static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
            if (m_buf.size() < 96 || std::memcmp(m_buf.data(), sig, 8) != 0 || m_buf[8] != 0 || m_buf[13] != 8 || m_buf[14] != 8)
            {
                throw std::runtime_error("'" + file + "' is not an HDF5 file this reader understands (superblock version 0, 8-byte offsets)");
            }
        }

        bool exists(const std::string& path) const
        {
            Entry e;
            return lookup(path, e);
        }

        template <class T>
        std::vector<T> read(const std::string& path, std::vector<uint64_t>* shape_out = nullptr) const
        {
            static_assert(sizeof(T) == 8, "64-bit element types only");
            Entry e;
            if (!lookup(path, e))
            {
                throw std::runtime_error("dataset '" + path + "' not found");
            }
            std::vector<uint64_t> shape;
            uint64_t addr = UNDEF, size = 0;
            bool have_layout = false;
            for_each_message(e.header,
                             [&](uint16_t type, uint64_t body, uint16_t)
                             {
                                 if (type == 0x0001)
                                 {
                                     const int ver = m_buf[body], rank = m_buf[body + 1];
                                     const uint64_t p = body + (ver == 1 ? 8 : 4);
                                     for (int i = 0; i < rank; ++i)
                                     {
                                         shape.push_back(get64(p + 8 * static_cast<uint64_t>(i)));
                                     }
                                 }
                                 else if (type == 0x0003)
                                 {
                                     if (get32(body + 4) != 8)
                                     {
                                         throw std::runtime_error("dataset '" + path + "': only 8-byte element types are supported");
                                     }
                                 }
                                 else if (type == 0x0008)
                                 {
                                     if (m_buf[body] != 3 || m_buf[body + 1] != 1)
                                     {
                                         throw std::runtime_error("dataset '" + path + "': only contiguous layout (version 3) is supported");
                                     }
                                     addr        = get64(body + 2);
                                     size        = get64(body + 10);
                                     have_layout = true;
                                 }
                             });
            if (!have_layout)
            {
                throw std::runtime_error("'" + path + "' is not a dataset");
            }
            uint64_t n = 1;
            for (uint64_t s : shape)
            {
                n *= s;
            }
            std::vector<T> out(n);
            if (addr != UNDEF && n > 0)
            {
                if (size < n * 8 || addr + n * 8 > m_buf.size())
                {
                    throw std::runtime_error("dataset '" + path + "' is truncated");
                }
                std::memcpy(out.data(), m_buf.data() + addr, n * 8);
            }
            if (shape_out)
            {
                *shape_out = shape;
            }
            return out;
        }

        std::vector<std::string> list(const std::string& path) const
        {
            Entry e;
            if (!lookup(path, e))
            {
                throw std::runtime_error("group '" + path + "' not found");
            }
            std::vector<std::string> names;
            for (auto& kv : group_entries(e))
            {
                names.push_back(kv.first);
            }
            return names;
        }

      private:

        struct Entry
        {
            uint64_t header = 0;
            uint32_t cache  = 0;
            uint64_t btree = 0, heap = 0;
        };

        uint64_t get64(uint64_t at) const
        {
            uint64_t v;
            std::memcpy(&v, m_buf.data() + at, 8);
            return v;
        }

        uint32_t get32(uint64_t at) const
        {
            uint32_t v;
            std::memcpy(&v, m_buf.data() + at, 4);
            return v;
        }

        uint16_t get16(uint64_t at) const
        {
            uint16_t v;
            std::memcpy(&v, m_buf.data() + at, 2);
            return v;
        }

        Entry entry_at(uint64_t at) const
        {
            Entry e;
            e.header = get64(at + 8);
            e.cache  = get32(at + 16);
            e.btree  = get64(at + 24);
            e.heap   = get64(at + 32);
            return e;
        }

        template <class F>
        void for_each_message(uint64_t header, F&& f) const
        {
            if (m_buf[header] != 1)
            {
                throw std::runtime_error("only version-1 object headers are supported");
            }
            const int nmsg = get16(header + 2);
            std::vector<std::pair<uint64_t, uint64_t>> blocks{{header + 16, get32(header + 8)}};
            int seen = 0;
            for (std::size_t bi = 0; bi < blocks.size() && seen < nmsg; ++bi)
            {
                uint64_t off = blocks[bi].first;
                const uint64_t end = off + blocks[bi].second;
                while (off + 8 <= end && seen < nmsg)
                {
                    const uint16_t type = get16(off), size = get16(off + 2);
                    if (type == 0x0010)
                    {
                        blocks.emplace_back(get64(off + 8), get64(off + 16));
                    }
                    f(type, off + 8, size);
                    off += 8 + size;
                    ++seen;
                }
            }
        }

        std::map<std::string, Entry> group_entries(Entry e) const
        {
            if (e.cache != 1)
            {
                bool found = false;
                for_each_message(e.header,
                                 [&](uint16_t type, uint64_t body, uint16_t)
                                 {
                                     if (type == 0x0011)
                                     {
                                         e.btree = get64(body);
                                         e.heap  = get64(body + 8);
                                         found   = true;
                                     }
                                 });
                if (!found)
                {
                    throw std::runtime_error("not a group");
                }
            }
            const uint64_t heap_data = get64(e.heap + 24);
            std::map<std::string, Entry> out;
            std::vector<uint64_t> stack{e.btree};
            while (!stack.empty())
            {
                const uint64_t node = stack.back();
                stack.pop_back();
                const int level = m_buf[node + 5], used = get16(node + 6);
                for (int i = 0; i < used; ++i)
                {
                    const uint64_t child = get64(node + 24 + 16 * static_cast<uint64_t>(i) + 8);
                    if (level > 0)
                    {
                        stack.push_back(child);
                    }
                    else
                    {
                        const int nsym = get16(child + 6);
                        for (int k = 0; k < nsym; ++k)
                        {
                            const uint64_t at = child + 8 + 40 * static_cast<uint64_t>(k);
                            const char* name  = reinterpret_cast<const char*>(m_buf.data() + heap_data + get64(at));
                            out[name]         = entry_at(at);
                        }
                    }
                }
            }
            return out;
        }

        bool lookup(const std::string& path, Entry& e) const
        {
            e = entry_at(56);
            std::size_t pos = 1;
            while (pos < path.size())
            {
                const std::size_t slash = path.find('/', pos);
                const std::string name  = path.substr(pos, slash == std::string::npos ? std::string::npos : slash - pos);
                auto entries            = group_entries(e);
                auto it                 = entries.find(name);
                if (it == entries.end())
                {
                    return false;
                }
                e = it->second;
                if (slash == std::string::npos)
                {
                    break;
                }
                pos = slash + 1;
            }
            return true;
        }

        std::vector<uint8_t> m_buf;
    };
} // namespace samurai::b200::h5
