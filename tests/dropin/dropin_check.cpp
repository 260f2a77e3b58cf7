// dropin_check.cpp -- runs the REFERENCE's own host code (gko::matrix::Csr::apply,
// gko::solver::{Cg,Bicgstab,Gmres}::build()...on(exec)->generate(A)->apply(b, x),
// gko::preconditioner::Jacobi, gko::stop::*) on a gko::CudaExecutor whose kernels are the B200
// library (ginkgo_b200/dropin/cuda_backend.cpp, built by tests/dropin/Makefile, linked in place of the reference's
// core/device_hooks/cuda_hooks.cpp stub) and compares every result with the same code on
// gko::ReferenceExecutor.  Written against the reference's public headers only.
//
//   dropin_check cuda            all cases on CudaExecutor vs ReferenceExecutor (GPU box)
//   dropin_check omp  [--trace]  the same flow on OmpExecutor (no GPU needed); --trace prints the
//                                name of every operation the flow launches = the list of
//                                gko::kernels::cuda::* symbols the backend has to define
// exit code 0 = all cases within tolerance.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include <ginkgo/core/base/array.hpp>
#include <ginkgo/core/base/executor.hpp>
#include <ginkgo/core/base/matrix_data.hpp>
#include <ginkgo/core/log/convergence.hpp>
#include <ginkgo/core/log/logger.hpp>
#include <ginkgo/core/matrix/coo.hpp>
#include <ginkgo/core/matrix/csr.hpp>
#include <ginkgo/core/matrix/dense.hpp>
#include <ginkgo/core/matrix/ell.hpp>
#include <ginkgo/core/matrix/hybrid.hpp>
#include <ginkgo/core/matrix/sellp.hpp>
#include <ginkgo/core/preconditioner/jacobi.hpp>
#include <ginkgo/core/solver/bicgstab.hpp>
#include <ginkgo/core/solver/cg.hpp>
#include <ginkgo/core/solver/gmres.hpp>
#include <ginkgo/core/stop/combined.hpp>
#include <ginkgo/core/stop/iteration.hpp>
#include <ginkgo/core/stop/residual_norm.hpp>

namespace {

struct OpTrace : gko::log::Logger {
    mutable std::map<std::string, long> count;
    OpTrace() : gko::log::Logger(gko::log::Logger::operation_launched_mask) {}
    void on_operation_launched(const gko::Executor*, const gko::Operation* op) const override
    {
        count[op->get_name()]++;
    }
};

template <typename V>
gko::matrix_data<V, gko::int32> laplace2d(int g)
{
    gko::matrix_data<V, gko::int32> d(gko::dim<2>(g * g, g * g));
    for (int i = 0; i < g; ++i)
        for (int j = 0; j < g; ++j) {
            const int r = i * g + j;
            if (i > 0) d.nonzeros.emplace_back(r, r - g, V(-1));
            if (j > 0) d.nonzeros.emplace_back(r, r - 1, V(-1));
            d.nonzeros.emplace_back(r, r, V(4));
            if (j < g - 1) d.nonzeros.emplace_back(r, r + 1, V(-1));
            if (i < g - 1) d.nonzeros.emplace_back(r, r + g, V(-1));
        }
    return d;
}

// nonsymmetric, row-diagonally dominant (BiCGStab / GMRES cases)
template <typename V>
gko::matrix_data<V, gko::int32> nonsym(int n, int per_row)
{
    gko::matrix_data<V, gko::int32> d(gko::dim<2>(n, n));
    unsigned long long s = 88172645463325252ull;
    auto rnd = [&] {
        s ^= s << 13;
        s ^= s >> 7;
        s ^= s << 17;
        return s;
    };
    for (int r = 0; r < n; ++r) {
        std::map<int, V> row;
        double sum = 0;
        while ((int)row.size() < per_row - 1) {
            const int c = (int)(rnd() % (unsigned long long)n);
            if (c == r || row.count(c)) continue;
            const double v = (double)(rnd() % 2000001ull) / 1e6 - 1.0;
            row[c] = (V)v;
            sum += std::abs((double)(V)v);
        }
        row[r] = (V)(sum + 1.0);
        for (auto& kv : row) d.nonzeros.emplace_back(r, kv.first, kv.second);
    }
    return d;
}

template <typename V>
double rel_diff(const gko::matrix::Dense<V>* a, const gko::matrix::Dense<V>* b)
{
    auto host = gko::ReferenceExecutor::create();
    auto ha = gko::clone(host, a);
    auto hb = gko::clone(host, b);
    double num = 0, den = 0;
    for (gko::size_type i = 0; i < ha->get_size()[0]; ++i)
        for (gko::size_type j = 0; j < ha->get_size()[1]; ++j) {
            const double x = ha->at(i, j), y = hb->at(i, j);
            num += (x - y) * (x - y);
            den += y * y;
        }
    return std::sqrt(num) / (den > 0 ? std::sqrt(den) : 1.0);
}

int failures = 0;
void report(const char* what, double err, double tol, long it_dev = -1, long it_ref = -1)
{
    // fp64: +-2 iterations; fp32 Krylov paths are chaotic w.r.t. the order of the dot-product sums
    // (tree on the device / threads, sequential on the reference): comparable count, same answer
    const long it_tol = tol > 1e-5 ? std::max(3L, it_ref / 4) : 2L;
    const bool it_ok = it_dev < 0 || std::labs(it_dev - it_ref) <= it_tol;
    const bool ok = err <= tol && it_ok;
    if (it_dev >= 0)
        std::printf("%-58s rel diff %.3e (tol %.0e)  iterations %ld vs reference %ld  %s\n", what, err, tol,
                    it_dev, it_ref, ok ? "OK" : "FAIL");
    else
        std::printf("%-58s rel diff %.3e (tol %.0e)  %s\n", what, err, tol, ok ? "OK" : "FAIL");
    if (!ok) ++failures;
}

template <typename V, typename Mtx>
void spmv_case(std::shared_ptr<gko::Executor> ref, std::shared_ptr<gko::Executor> dev,
               const gko::matrix_data<V, gko::int32>& data, const char* name, int nrhs)
{
    using Dense = gko::matrix::Dense<V>;
    const auto n = data.size[0];
    auto A_ref = Mtx::create(ref);
    A_ref->read(data);
    auto A_dev = gko::clone(dev, A_ref);
    auto b = Dense::create(ref, gko::dim<2>(data.size[1], nrhs));
    for (gko::size_type i = 0; i < data.size[1]; ++i)
        for (int j = 0; j < nrhs; ++j) b->at(i, j) = V(std::sin(0.37 * (double)i + j));
    auto x_ref = Dense::create(ref, gko::dim<2>(n, nrhs));
    x_ref->fill(V(0.5));
    auto b_dev = gko::clone(dev, b);
    auto x_dev = gko::clone(dev, x_ref);
    const double tol = sizeof(V) == 8 ? 1e-14 : 1e-6;
    char what[128];
    A_ref->apply(b, x_ref);
    A_dev->apply(b_dev, x_dev);
    std::snprintf(what, sizeof what, "%s::apply  (%s, %d rhs)", name, sizeof(V) == 8 ? "f64" : "f32", nrhs);
    report(what, rel_diff(x_dev.get(), x_ref.get()), tol);
    auto alpha = gko::initialize<Dense>({V(-1.5)}, ref);
    auto beta = gko::initialize<Dense>({V(0.25)}, ref);
    A_ref->apply(alpha, b, beta, x_ref);
    A_dev->apply(gko::clone(dev, alpha), b_dev, gko::clone(dev, beta), x_dev);
    std::snprintf(what, sizeof what, "%s::apply(alpha, b, beta, x)  (%s, %d rhs)", name,
                  sizeof(V) == 8 ? "f64" : "f32", nrhs);
    report(what, rel_diff(x_dev.get(), x_ref.get()), tol);
}

template <typename V, typename Solver>
void solver_case(std::shared_ptr<gko::Executor> ref, std::shared_ptr<gko::Executor> dev,
                 const gko::matrix_data<V, gko::int32>& data, const char* name, unsigned max_block_size,
                 double reduction)
{
    using Dense = gko::matrix::Dense<V>;
    using Csr = gko::matrix::Csr<V, gko::int32>;
    using Jacobi = gko::preconditioner::Jacobi<V, gko::int32>;
    const auto n = data.size[0];
    auto run = [&](std::shared_ptr<gko::Executor> exec, std::unique_ptr<Dense>& x, long& iters) {
        // assembled on the host, moved to the executor: the device side of Csr::read
        // (device_matrix_data, aos_to_soa, sort) is set-up outside the hot path
        auto A_host = Csr::create(exec->get_master());
        A_host->read(data);
        auto A = gko::share(gko::clone(exec, A_host));
        auto b = Dense::create(exec, gko::dim<2>(n, 1));
        b->fill(V(1));
        x = Dense::create(exec, gko::dim<2>(n, 1));
        x->fill(V(0));
        auto logger = gko::share(gko::log::Convergence<V>::create());
        auto iter_stop = gko::share(gko::stop::Iteration::build().with_max_iters(2000u).on(exec));
        auto res_stop = gko::share(gko::stop::ResidualNorm<V>::build()
                                       .with_baseline(gko::stop::mode::rhs_norm)
                                       .with_reduction_factor((gko::remove_complex<V>)reduction)
                                       .on(exec));
        std::shared_ptr<gko::LinOp> solver;
        if (max_block_size > 0) {
            solver = Solver::build()
                         .with_criteria(iter_stop, res_stop)
                         .with_preconditioner(Jacobi::build().with_max_block_size(max_block_size))
                         .on(exec)
                         ->generate(A);
        } else {
            solver = Solver::build().with_criteria(iter_stop, res_stop).on(exec)->generate(A);
        }
        solver->add_logger(logger);
        solver->apply(b, x);
        exec->synchronize();
        iters = (long)logger->get_num_iterations();
    };
    std::unique_ptr<Dense> x_ref, x_dev;
    long it_ref = 0, it_dev = 0;
    run(ref, x_ref, it_ref);
    run(dev, x_dev, it_dev);
    char what[160];
    std::snprintf(what, sizeof what, "solver::%s + Jacobi(max_block_size=%u) (%s, n=%lld)", name, max_block_size,
                  sizeof(V) == 8 ? "f64" : "f32", (long long)n);
    report(what, rel_diff(x_dev.get(), x_ref.get()), sizeof(V) == 8 ? 1e-7 : 2e-3, it_dev, it_ref);
}

// the reference's own adaptive-precision block-Jacobi (storage_optimization = autodetect) generated and
// applied on the device executor: chosen precisions and condition numbers must be IDENTICAL to the
// ReferenceExecutor's, the stored blocks bit-identical, the apply within tolerance
template <typename V>
void adaptive_jacobi_case(std::shared_ptr<gko::Executor> ref, std::shared_ptr<gko::Executor> dev,
                          const gko::matrix_data<V, gko::int32>& data, unsigned max_block_size, bool transposed)
{
    using Dense = gko::matrix::Dense<V>;
    using Csr = gko::matrix::Csr<V, gko::int32>;
    using Jacobi = gko::preconditioner::Jacobi<V, gko::int32>;
    const auto n = data.size[0];
    struct Out {
        std::unique_ptr<Dense> x;
        std::vector<unsigned char> prec, bytes;
        std::vector<double> cond;
    };
    auto run = [&](std::shared_ptr<gko::Executor> exec, Out& o) {
        auto A_host = Csr::create(exec->get_master());
        A_host->read(data);
        auto A = gko::share(gko::clone(exec, A_host));
        gko::array<gko::int32> ptrs(exec->get_master(), (n + max_block_size - 1) / max_block_size + 1);
        for (gko::size_type k = 0; k < ptrs.get_size(); ++k)
            ptrs.get_data()[k] = (gko::int32)std::min<gko::size_type>(k * max_block_size, n);
        std::shared_ptr<Jacobi> J = Jacobi::build()
                                        .with_max_block_size(max_block_size)
                                        .with_block_pointers(gko::array<gko::int32>(exec, ptrs))
                                        .with_storage_optimization(gko::precision_reduction::autodetect())
                                        .with_accuracy((gko::remove_complex<V>)0.1)
                                        .on(exec)
                                        ->generate(A);
        if (transposed) J = gko::as<Jacobi>(J->transpose());
        auto b = Dense::create(exec, gko::dim<2>(n, 2));
        b->fill(V(1));
        o.x = Dense::create(exec, gko::dim<2>(n, 2));
        o.x->fill(V(0));
        J->apply(b, o.x);
        exec->synchronize();
        const auto nb = J->get_num_blocks();
        gko::array<gko::precision_reduction> pr(exec->get_master(), J->get_parameters().storage_optimization.block_wise);
        for (gko::size_type k = 0; k < nb; ++k) o.prec.push_back((unsigned char)pr.get_const_data()[k]);
        gko::array<gko::remove_complex<V>> cond(exec, nb);
        exec->copy(nb, J->get_conditioning(), cond.get_data());
        cond.set_executor(exec->get_master());
        for (gko::size_type k = 0; k < nb; ++k) o.cond.push_back((double)cond.get_const_data()[k]);
        // stored bits of the blocks actually written: position (r, c) of block k in its precision
        gko::array<V> blocks(exec, J->get_num_stored_elements());
        exec->copy(J->get_num_stored_elements(), J->get_blocks(), blocks.get_data());
        blocks.set_executor(exec->get_master());
        const auto sch = J->get_storage_scheme();
        for (gko::size_type k = 0; k < nb; ++k) {
            const auto p = o.prec[k];
            const int w = sizeof(V) == 8 ? (p == 0x01 || p == 0x10 ? 4 : (p == 0x00 ? 8 : 2)) : (p == 0x00 ? 4 : 2);
            const auto* base = reinterpret_cast<const unsigned char*>(blocks.get_const_data() + sch.get_group_offset(k));
            const int bs = ptrs.get_const_data()[k + 1] - ptrs.get_const_data()[k];
            for (int c = 0; c < bs; ++c)
                for (int r = 0; r < bs; ++r) {
                    const std::size_t idx = sch.get_block_offset(k) + r + c * sch.get_stride();
                    for (int q = 0; q < w; ++q) o.bytes.push_back(base[idx * w + q]);
                }
        }
    };
    Out a, d;
    run(ref, a);
    run(dev, d);
    char what[160];
    std::snprintf(what, sizeof what, "preconditioner::Jacobi(%u, autodetect)%s chosen precisions (%s)", max_block_size,
                  transposed ? "^T" : "", sizeof(V) == 8 ? "f64" : "f32");
    report(what, a.prec == d.prec ? 0.0 : 1.0, 0.0);
    std::snprintf(what, sizeof what, "preconditioner::Jacobi(%u, autodetect)%s condition numbers", max_block_size,
                  transposed ? "^T" : "");
    report(what, a.cond == d.cond ? 0.0 : 1.0, 0.0);
    std::snprintf(what, sizeof what, "preconditioner::Jacobi(%u, autodetect)%s stored block bits", max_block_size,
                  transposed ? "^T" : "");
    report(what, a.bytes == d.bytes ? 0.0 : 1.0, 0.0);
    std::snprintf(what, sizeof what, "preconditioner::Jacobi(%u, autodetect)%s::apply (2 rhs)", max_block_size,
                  transposed ? "^T" : "");
    report(what, rel_diff(d.x.get(), a.x.get()), sizeof(V) == 8 ? 1e-14 : 1e-6);
    int hist[256] = {};
    for (auto p : d.prec) hist[p]++;
    std::printf("#   precisions on the device: (0,0) %d  (0,1) %d  (0,2) %d  (1,0) %d  (1,1) %d  (2,0) %d\n", hist[0x00],
                hist[0x01], hist[0x02], hist[0x10], hist[0x11], hist[0x20]);
}

}  // namespace

int main(int argc, char** argv)
{
    const std::string which = argc > 1 ? argv[1] : "cuda";
    const bool trace = argc > 2 && !std::strcmp(argv[2], "--trace");
    auto ref = gko::ReferenceExecutor::create();
    std::shared_ptr<gko::Executor> dev;
    try {
        if (which == "cuda")
            dev = gko::CudaExecutor::create(0, ref);
        else
            dev = gko::OmpExecutor::create();
        std::printf("# device executor: %s\n", dev->get_description().c_str());
    } catch (const std::exception& e) {
        std::printf("cannot create the %s executor: %s\n", which.c_str(), e.what());
        return 3;
    }
    auto tr = std::make_shared<OpTrace>();
    if (trace) dev->add_logger(tr);
    try {
        auto lap = laplace2d<double>(48);
        auto lap_f = laplace2d<float>(48);
        auto ns = nonsym<double>(3000, 9);
        auto ns_f = nonsym<float>(3000, 9);
        spmv_case<double, gko::matrix::Csr<double, gko::int32>>(ref, dev, ns, "matrix::Csr", 1);
        spmv_case<float, gko::matrix::Csr<float, gko::int32>>(ref, dev, ns_f, "matrix::Csr", 1);
        spmv_case<double, gko::matrix::Csr<double, gko::int32>>(ref, dev, ns, "matrix::Csr", 3);
        spmv_case<double, gko::matrix::Ell<double, gko::int32>>(ref, dev, ns, "matrix::Ell", 1);
        spmv_case<double, gko::matrix::Sellp<double, gko::int32>>(ref, dev, ns, "matrix::Sellp", 1);
        spmv_case<double, gko::matrix::Coo<double, gko::int32>>(ref, dev, ns, "matrix::Coo", 1);
        spmv_case<double, gko::matrix::Hybrid<double, gko::int32>>(ref, dev, ns, "matrix::Hybrid", 1);
        solver_case<double, gko::solver::Cg<double>>(ref, dev, lap, "Cg", 0, 1e-10);
        solver_case<double, gko::solver::Cg<double>>(ref, dev, lap, "Cg", 1, 1e-10);
        solver_case<double, gko::solver::Cg<double>>(ref, dev, lap, "Cg", 8, 1e-10);
        solver_case<float, gko::solver::Cg<float>>(ref, dev, lap_f, "Cg", 1, 1e-5);
        solver_case<double, gko::solver::Bicgstab<double>>(ref, dev, ns, "Bicgstab", 1, 1e-10);
        solver_case<double, gko::solver::Gmres<double>>(ref, dev, ns, "Gmres", 1, 1e-10);
        solver_case<float, gko::solver::Gmres<float>>(ref, dev, ns_f, "Gmres", 8, 1e-5);
        adaptive_jacobi_case<double>(ref, dev, ns, 16, false);
        adaptive_jacobi_case<double>(ref, dev, lap, 7, false);
        adaptive_jacobi_case<double>(ref, dev, ns, 32, true);
        adaptive_jacobi_case<float>(ref, dev, ns_f, 16, false);
    } catch (const std::exception& e) {
        std::printf("EXCEPTION: %s\n", e.what());
        ++failures;
    }
    if (trace) {
        std::printf("# operations launched on the device executor (name : count)\n");
        for (auto& kv : tr->count) std::printf("OP %s : %ld\n", kv.first.c_str(), kv.second);
    }
    std::printf("%s (%d failing case%s)\n", failures ? "DROPIN CHECK FAILED" : "DROPIN CHECK PASSED", failures,
                failures == 1 ? "" : "s");
    return failures ? 1 : 0;
}
