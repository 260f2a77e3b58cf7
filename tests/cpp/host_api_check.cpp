// TEST INFRASTRUCTURE: the reference's test idioms on the host layer (compiled against the
// host-memory mock by tests/test_host_cpu.py): gko::initialize / share / clone, the stencil solves
// of reference/test/solver/{cg,bicgstab,gmres}_kernels.cpp and a format round trip.
#include <chrono>
#include <cmath>
#include <cstdio>

#include "../../ginkgo_b200/host/gko_b200.hpp"

namespace gko = gko_b200;

template <typename Solver>
static int solve_stencil(std::shared_ptr<const gko::Executor> exec, const char* name)
{
    using Mtx = gko::matrix::Dense<double>;
    using Csr = gko::matrix::Csr<double, gko::int32>;
    auto mtx = gko::share(gko::initialize<Csr>({{2, -1.0, 0.0}, {-1.0, 2, -1.0}, {0.0, -1.0, 2}}, exec));
    auto factory = Solver::build()
                       .with_criteria(gko::stop::Iteration::build().with_max_iters(4u),
                                      gko::stop::Time::build().with_time_limit(std::chrono::seconds(6)),
                                      gko::stop::ResidualNorm<double>::build().with_reduction_factor(1e-14))
                       .on(exec);
    auto solver = factory->generate(mtx);
    auto b = gko::initialize<Mtx>({-1.0, 3.0, 1.0}, exec);
    auto x = gko::initialize<Mtx>({0.0, 0.0, 0.0}, exec);
    solver->apply(b, x);
    const auto h = x->to_host();
    const double want[3] = {1.0, 3.0, 2.0};
    double err = 0;
    for (int i = 0; i < 3; ++i) err = std::fmax(err, std::fabs(h[i] - want[i]));
    std::printf("%s stencil: %.15g %.15g %.15g\n", name, h[0], h[1], h[2]);
    return err < 1e-13 ? 0 : 1;
}

int main()
{
    using Mtx = gko::matrix::Dense<double>;
    auto exec = gko::B200Executor::create(0);
    int bad = 0;
    bad += solve_stencil<gko::solver::Cg<double>>(exec, "cg");
    bad += solve_stencil<gko::solver::Bicgstab<double>>(exec, "bicgstab");
    bad += solve_stencil<gko::solver::Gmres<double>>(exec, "gmres");
    bad += solve_stencil<gko::solver::Bicg<double>>(exec, "bicg");

    // initialize<Ell / Sellp / Coo / Hybrid>, apply, clone: reference/test/matrix/csr_kernels.cpp:84-106
    auto x = gko::initialize<Mtx>({2.0, 1.0, 4.0}, exec);
    auto check = [&](const gko::LinOp* a, const char* name) {
        auto y = Mtx::create(exec, gko::dim2{2, 1});
        a->apply(x.get(), y.get());
        auto yc = gko::clone(y);
        const auto h = yc->to_host();
        std::printf("%s apply: %g %g\n", name, h[0], h[1]);
        return (h[0] == 13.0 && h[1] == 5.0) ? 0 : 1;
    };
    bad += check(gko::initialize<gko::matrix::Csr<double, gko::int32>>({{1.0, 3.0, 2.0}, {0.0, 5.0, 0.0}}, exec).get(), "csr");
    bad += check(gko::initialize<gko::matrix::Ell<double, gko::int32>>({{1.0, 3.0, 2.0}, {0.0, 5.0, 0.0}}, exec).get(), "ell");
    bad += check(gko::initialize<gko::matrix::Sellp<double, gko::int32>>({{1.0, 3.0, 2.0}, {0.0, 5.0, 0.0}}, exec).get(), "sellp");
    bad += check(gko::initialize<gko::matrix::Coo<double, gko::int32>>({{1.0, 3.0, 2.0}, {0.0, 5.0, 0.0}}, exec).get(), "coo");
    bad += check(gko::initialize<gko::matrix::Hybrid<double, gko::int32>>({{1.0, 3.0, 2.0}, {0.0, 5.0, 0.0}}, exec).get(), "hybrid");
    {  // stop::Time with an expired limit: the solve stops before the first iteration, criterion id 1
        using Csr = gko::matrix::Csr<double, gko::int32>;
        auto mtx = gko::share(gko::initialize<Csr>({{2, -1.0, 0.0}, {-1.0, 2, -1.0}, {0.0, -1.0, 2}}, exec));
        auto solver = gko::solver::Cg<double>::build()
                          .with_criteria(gko::stop::Time::build().with_time_limit(std::chrono::nanoseconds(0)),
                                         gko::stop::Iteration::build().with_max_iters(100u))
                          .on(exec)
                          ->generate(mtx);
        auto b = gko::initialize<Mtx>({-1.0, 3.0, 1.0}, exec);
        auto xs = gko::initialize<Mtx>({0.0, 0.0, 0.0}, exec);
        solver->apply(b, xs);
        auto base = dynamic_cast<gko::solver::SolverBase<double>*>(solver.get());
        std::printf("time criterion: iterations %zu status %#x\n", (size_t)base->get_num_iterations(),
                    (unsigned)base->get_stop_status());
        bad += !(base->get_num_iterations() == 0 && (base->get_stop_status() & 0x3f) == 1 &&
                 !base->has_converged());
    }
    auto dense = gko::initialize<Mtx>({{1.0, 2.0}, {3.0, 4.0}}, exec);
    const auto dh = dense->to_host();
    bad += !(dense->get_size().rows == 2 && dense->get_size().cols == 2 && dh[2] == 3.0);
    std::printf("%s\n", bad ? "FAILED" : "ALL OK");
    return bad;
}
