// The reference's examples/distributed-solver/distributed-solver.cpp flow on the B200 host
// layer: one process per GPU, a uniform row partition, Matrix::read_distributed of the global
// stencil matrix, distributed vectors, a Krylov solver with a Jacobi preconditioner generated
// from the local block.  MPI is replaced by the library's own communicator (NCCL / peer memory):
//   RANK, WORLD_SIZE, LOCAL_RANK   from the launcher (torchrun, srun, mpirun wrapper ...)
//   B200_ID_FILE                   a path all ranks can see; rank 0 drops the 128-byte unique id there
//   g++ -O2 -std=c++17 examples/distributed_solver.cpp -Lginkgo_b200/lib -lginkgo_b200 -o distributed_solver
//   usage: distributed_solver [grid = 64] [solver = cg | gmres | bicgstab]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

#include "../ginkgo_b200/host/gko_b200_dist.hpp"

namespace gko = gko_b200;

static int env_int(const char* name, int fallback)
{
    const char* v = std::getenv(name);
    return v ? std::atoi(v) : fallback;
}

int main(int argc, char** argv)
{
    using ValueType = double;
    using LocalIndexType = gko::int32;
    using GlobalIndexType = gko::int64;
    using dist_mtx = gko::distributed::Matrix<ValueType, LocalIndexType>;
    using dist_vec = gko::distributed::Vector<ValueType>;
    using part_type = gko::distributed::Partition<LocalIndexType, GlobalIndexType>;
    using bj = gko::preconditioner::Jacobi<ValueType, LocalIndexType>;

    const int rank = env_int("RANK", 0), size = env_int("WORLD_SIZE", 1);
    const int grid = argc > 1 ? std::atoi(argv[1]) : 64;
    const std::string solver_name = argc > 2 ? argv[2] : "cg";
    const auto exec = gko::B200Executor::create(env_int("LOCAL_RANK", rank));

    // the unique id of the communicator: created by rank 0, read by the others
    gko::uint8 id[128] = {};
    const char* id_file = std::getenv("B200_ID_FILE");
    if (size > 1 && !id_file) {
        std::fprintf(stderr, "set B200_ID_FILE to a path all ranks can see\n");
        return 2;
    }
    if (rank == 0) {
        gko::distributed::communicator::get_unique_id(id);
        if (size > 1) {
            std::ofstream(std::string(id_file) + ".tmp", std::ios::binary).write((const char*)id, 128);
            std::rename((std::string(id_file) + ".tmp").c_str(), id_file);
        }
    } else {
        for (int tries = 0; tries < 600; ++tries) {
            std::ifstream in(id_file, std::ios::binary);
            if (in.read((char*)id, 128)) break;
            std::this_thread::sleep_for(std::chrono::milliseconds(100));
        }
    }
    auto comm = gko::distributed::communicator::create(exec, id, rank, size);

    // 7-point Laplacian on a grid^3 box; every rank assembles only the rows it owns
    const GlobalIndexType n = (GlobalIndexType)grid * grid * grid;
    auto partition = std::shared_ptr<const part_type>(part_type::build_from_global_size_uniform(exec, size, n));
    const auto bounds = gko::array<GlobalIndexType>::view(exec, size + 1,
                                                          const_cast<GlobalIndexType*>(partition->get_range_bounds()))
                            .to_host();
    gko::matrix_data<ValueType, GlobalIndexType> data(gko::dim2{(gko::size_type)n, (gko::size_type)n});
    for (GlobalIndexType r = bounds[rank]; r < bounds[rank + 1]; ++r) {
        const GlobalIndexType x = r % grid, y = (r / grid) % grid, z = r / ((GlobalIndexType)grid * grid);
        if (z > 0) data.nonzeros.push_back({r, r - (GlobalIndexType)grid * grid, -1.0});
        if (y > 0) data.nonzeros.push_back({r, r - grid, -1.0});
        if (x > 0) data.nonzeros.push_back({r, r - 1, -1.0});
        data.nonzeros.push_back({r, r, 6.0});
        if (x < grid - 1) data.nonzeros.push_back({r, r + 1, -1.0});
        if (y < grid - 1) data.nonzeros.push_back({r, r + grid, -1.0});
        if (z < grid - 1) data.nonzeros.push_back({r, r + (GlobalIndexType)grid * grid, -1.0});
    }
    auto A = dist_mtx::read_distributed<GlobalIndexType>(exec, comm, data, partition);

    const gko::size_type n_local = A->n_local();
    auto b = dist_vec::create(exec, comm, gko::dim2{(gko::size_type)n, 1}, gko::dim2{n_local, 1});
    auto x = dist_vec::create(exec, comm, gko::dim2{(gko::size_type)n, 1}, gko::dim2{n_local, 1});
    b->fill(1.0);
    x->fill(0.0);

    auto criteria_iter = gko::stop::Iteration::build().with_max_iters(2000u);
    auto criteria_res = gko::stop::ResidualNorm<ValueType>::build().with_reduction_factor(1e-8);
    auto precond = bj::build().with_max_block_size(1u).on(exec);
    std::unique_ptr<gko::LinOp> solver;
    if (solver_name == "gmres")
        solver = gko::solver::Gmres<ValueType>::build()
                     .with_criteria(criteria_iter, criteria_res)
                     .with_preconditioner(precond)
                     .on(exec)
                     ->generate(A);
    else if (solver_name == "bicgstab")
        solver = gko::solver::Bicgstab<ValueType>::build()
                     .with_criteria(criteria_iter, criteria_res)
                     .with_preconditioner(precond)
                     .on(exec)
                     ->generate(A);
    else
        solver = gko::solver::Cg<ValueType>::build()
                     .with_criteria(criteria_iter, criteria_res)
                     .with_preconditioner(precond)
                     .on(exec)
                     ->generate(A);
    exec->synchronize();
    const auto t0 = std::chrono::steady_clock::now();
    solver->apply(b.get(), x.get());
    exec->synchronize();
    const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    // ||b - A x|| / ||b||, both norms summed over the ranks by the vectors themselves
    auto one = gko::matrix::scalar<ValueType>(1.0, exec);
    auto neg_one = gko::matrix::scalar<ValueType>(-1.0, exec);
    auto res = gko::matrix::Dense<ValueType>::create(exec, gko::dim2{1, 1});
    auto bn = gko::matrix::Dense<ValueType>::create(exec, gko::dim2{1, 1});
    b->compute_norm2(bn.get());
    A->apply(neg_one.get(), x.get(), one.get(), b.get());
    b->compute_norm2(res.get());
    const double rel = res->to_host()[0] / bn->to_host()[0];
    auto base = dynamic_cast<gko::solver::SolverBase<ValueType>*>(solver.get());
    if (rank == 0)
        std::printf("ranks=%d n=%lld local_rows=%zu ghosts=%zu solver=%s iterations=%zu converged=%d "
                    "rel_residual=%.3e time=%.3fs\n",
                    size, (long long)n, (size_t)n_local, (size_t)A->n_ghost(), solver_name.c_str(),
                    (size_t)base->get_num_iterations(), (int)base->has_converged(), rel, seconds);
    return base->has_converged() && rel < 1e-7 ? 0 : 1;
}
