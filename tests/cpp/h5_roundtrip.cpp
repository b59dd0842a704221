// CPU-only check of include/samurai/b200_h5.hpp: writes a small file with the reference's save() layout and reads it back.
#include <samurai/b200_h5.hpp>

#include <cstdio>
#include <cstdlib>

int main(int argc, char** argv)
{
    using namespace samurai::b200::h5;
    const std::string file = argc > 1 ? argv[1] : "roundtrip.h5";
    std::vector<double> pts{0, 0, 0, 0.5, 0, 0, 0.5, 0.5, 0, 0, 0.5, 0, 1, 0, 0, 1, 0.5, 0};
    std::vector<uint64_t> conn{0, 1, 2, 3, 1, 4, 5, 2};
    std::vector<double> u{1.25, -3.5};
    std::vector<uint64_t> level{1, 1};
    std::vector<int64_t> ivl{1, 0, 0, 0, 2};
    Writer w;
    w.add("/mesh/connectivity", Type::u64, {2, 4}, conn.data());
    w.add("/mesh/points", Type::f64, {6, 3}, pts.data());
    w.add("/mesh/fields/u", Type::f64, {2}, u.data());
    w.add("/mesh/fields/level", Type::u64, {2}, level.data());
    w.add("/restart/intervals", Type::i64, {1, 5}, ivl.data());
    w.add_scalar("/n_process", uint64_t(1));
    w.add_scalar("/mesh/scaling_factor", 0.5);
    w.write(file);
    Reader r(file);
    std::vector<uint64_t> shape;
    bool ok = r.read<double>("/mesh/points", &shape) == pts && shape == std::vector<uint64_t>{6, 3};
    ok      = ok && r.read<uint64_t>("/mesh/connectivity") == conn && r.read<double>("/mesh/fields/u") == u && r.read<uint64_t>("/mesh/fields/level") == level;
    ok      = ok && r.read<int64_t>("/restart/intervals") == ivl && r.read<uint64_t>("/n_process").at(0) == 1 && r.read<double>("/mesh/scaling_factor").at(0) == 0.5;
    ok      = ok && r.exists("/mesh/fields") && !r.exists("/mesh/nope") && r.list("/mesh/fields") == std::vector<std::string>{"level", "u"};
    std::printf("%s\n", ok ? "roundtrip OK" : "roundtrip FAILED");
    return ok ? 0 : 1;
}
