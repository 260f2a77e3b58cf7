// The reference's examples/simple-solver/simple-solver.cpp flow on the B200 host layer:
// same builder chain, same apply calls (`namespace gko = gko_b200;`).
//   g++ -O2 -std=c++17 examples/simple_solver.cpp -Lginkgo_b200/lib -lginkgo_b200 -o simple_solver
#include <cstdio>
#include <vector>

#include "../ginkgo_b200/host/gko_b200.hpp"

namespace gko = gko_b200;

int main(int argc, char** argv)
{
    using ValueType = double;
    using IndexType = int;
    using vec = gko::matrix::Dense<ValueType>;
    using mtx = gko::matrix::Csr<ValueType, IndexType>;
    using cg = gko::solver::Cg<ValueType>;
    using bj = gko::preconditioner::Jacobi<ValueType, IndexType>;

    const int g = argc > 1 ? std::atoi(argv[1]) : 48;  // 7-pt Laplacian on a g^3 grid
    const gko::size_type n = (gko::size_type)g * g * g;
    std::vector<IndexType> rp(n + 1, 0), ci;
    std::vector<ValueType> va;
    for (int z = 0; z < g; ++z)
        for (int y = 0; y < g; ++y)
            for (int x = 0; x < g; ++x) {
                const gko::size_type r = ((gko::size_type)z * g + y) * g + x;
                auto add = [&](long c, double v) {
                    ci.push_back((IndexType)c);
                    va.push_back(v);
                };
                if (z > 0) add(r - (long)g * g, -1);
                if (y > 0) add(r - g, -1);
                if (x > 0) add(r - 1, -1);
                add(r, 6);
                if (x < g - 1) add(r + 1, -1);
                if (y < g - 1) add(r + g, -1);
                if (z < g - 1) add(r + (long)g * g, -1);
                rp[r + 1] = (IndexType)ci.size();
            }

    const auto exec = gko::B200Executor::create(0);
    auto A = std::shared_ptr<mtx>(mtx::create_from_host(exec, gko::dim2{n, n}, va, ci, rp));
    std::vector<ValueType> ones(n, 1.0), zeros(n, 0.0);
    auto b = vec::create_from_host(exec, gko::dim2{n, 1}, ones.data());
    auto x = vec::create_from_host(exec, gko::dim2{n, 1}, zeros.data());

    const double reduction_factor = 1e-8;
    auto solver_gen =
        cg::build()
            .with_criteria(gko::stop::Iteration::build().with_max_iters(2000u),
                           gko::stop::ResidualNorm<ValueType>::build().with_reduction_factor(
                               reduction_factor))
            .with_preconditioner(bj::build().with_max_block_size(1u).on(exec))
            .on(exec);
    auto solver = solver_gen->generate(A);
    solver->apply(b, x);

    // res = ||b - A x||
    auto one = gko::matrix::scalar<ValueType>(1.0, exec);
    auto neg_one = gko::matrix::scalar<ValueType>(-1.0, exec);
    auto res = vec::create(exec, gko::dim2{1, 1});
    auto bn = vec::create(exec, gko::dim2{1, 1});
    b->compute_norm2(bn.get());
    A->apply(neg_one, x, one, b);
    b->compute_norm2(res.get());
    auto s = dynamic_cast<cg*>(solver.get());
    std::printf("n=%zu iterations=%zu fused=%d converged=%d rel_residual=%.3e\n", (size_t)n,
                (size_t)s->get_num_iterations(), (int)s->used_fused_path(), (int)s->has_converged(),
                res->to_host()[0] / bn->to_host()[0]);
    return s->has_converged() ? 0 : 1;
}
