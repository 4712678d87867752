# baseline/run_reference.jl -- the reference's own CPU path on the BASELINE configs (SURVEY.md 8d), for anyone
# with Julia: its numbers supersede the restated NumPy baseline that bench.py --impl reference reports here
# (Julia is not installed in this project's build container or on its GPU boxes, so this script is shipped
# UNEXERCISED).
#
#   julia baseline/run_reference.jl [config] [workers]
#     config  : c1 | c2 | c3 | c4 (default c2: NIW N=1e6, D=32, K=20, alpha=10, 100 iterations)
#     workers : Distributed worker processes (default: all physical cores), one BLAS thread each
#               (README.md:43, docs/src/perf.md:6)
using Distributed
config = length(ARGS) >= 1 ? ARGS[1] : "c2"
nworkers_req = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : Sys.CPU_THREADS ÷ 2
addprocs(nworkers_req)
@everywhere using DPMMSubClusters
@everywhere using LinearAlgebra
@everywhere BLAS.set_num_threads(1)
using Random

cfg = Dict("c1" => (:niw, 10^4, 2, 6), "c2" => (:niw, 10^6, 32, 20), "c3" => (:mnm, 10^6, 100, 20), "c4" => (:niw, 10^7, 5, 50))[config]
kind, N, D, K = cfg
Random.seed!(0)
if kind == :niw
    x, labels, clusters = generate_gaussian_data(N, D, K, 100.0)
    hyper = DPMMSubClusters.niw_hyperparams(1.0, zeros(D), D + 3, Matrix{Float64}(I, D, D) * 1.0)
else
    x, labels, clusters = generate_mnmm_data(N, D, K, 50)
    hyper = DPMMSubClusters.multinomial_hyper(ones(Float32, D))
end
fit(x[:, 1:min(N, 10^4)], hyper, 10.0, iters = 3, verbose = false)          # compile
t = @elapsed ret = fit(x, hyper, 10.0, iters = 100, seed = 1, verbose = false, gt = labels, burnout = 20)
iter_times = ret[4]
println("""{"impl": "reference-julia", "config": "$config", "workers": $(nworkers()), "blas_threads": 1, """ *
        """"iters": 100, "seconds": $t, "iters_per_s": $(100 / t), "loop_iters_per_s": $(100 / sum(iter_times)), """ *
        """"final_K": $(length(ret[2])), "nmi": $(ret[5][end])}""")
