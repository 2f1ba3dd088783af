// bench_er -- stand-alone C++ driver with the reference's command line
// (bench_erdos_renyi.cpp:23-28):   bench_er logM edgeFactor algorithm R c output_file [trials [warmup]]
// algorithm: "15d" (15d_fusion1 + 15d_fusion2, as the reference), "25d" (25d_sparse_replicate unfused +
// 25d_dense_replicate fused), or any single selector of benchmark_algorithm ("15d_fusion1", "15d_fusion2",
// "15d_sparse", "25d_dense_replicate", "25d_sparse_replicate").
// One process per GPU: RANK / WORLD_SIZE / LOCAL_RANK in the environment, HNH_NCCL_ID_FILE when WORLD_SIZE > 1.
#include <cstdlib>
#include <iostream>
#include <string>

#include "hnh/benchmark_dist.hpp"

int main(int argc, char **argv) {
    if (argc < 7) {
        std::cerr << "usage: bench_er logM edgeFactor algorithm R c output_file [trials [warmup]]" << std::endl;
        return 2;
    }
    const int logM = atoi(argv[1]), edgeFactor = atoi(argv[2]);
    const std::string algorithm_name(argv[3]);
    const int R = atoi(argv[4]), c = atoi(argv[5]);
    const std::string output_file(argv[6]);
    const int trials = argc > 7 ? atoi(argv[7]) : 5, warmup = argc > 8 ? atoi(argv[8]) : 0;
    try {
        hnh_world_init_from_env();
        {
            SpmatLocal S;
            S.loadTuples(false, logM, edgeFactor, "");
            auto run = [&](const std::string &alg, bool fused) {
                json j = benchmark_algorithm_ex(&S, alg, output_file, fused, R, c, "vanilla", trials, warmup);
                if (hnh::Comm::world()->rank() == 0) std::cout << j.dump(2) << std::endl;
            };
            if (algorithm_name == "15d") {
                run("15d_fusion1", true);
                run("15d_fusion2", true);
            } else if (algorithm_name == "25d") {
                run("25d_sparse_replicate", false);
                run("25d_dense_replicate", true);
            } else {
                run(algorithm_name, true);
            }
        }
        hnh_world_finalize();
    } catch (const std::exception &e) {
        std::cerr << "bench_er: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
