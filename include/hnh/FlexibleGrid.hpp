// hnh/FlexibleGrid.hpp -- the 3-D process grid of the reference (FlexibleGrid.hpp:12-202)
// on hnh::Comm sub-communicators instead of MPI_Comm_split.
//
// (i, j, k) <-> world rank under the six "adjacency" orderings (reference :31-40,54-73,
// 105-135); row_world = fixed (i,k), ranks ordered by j; col_world = fixed (j,k), ordered by i;
// fiber_world = fixed (i,j), ordered by k (reference :80-82), hence rankInRow == j,
// rankInCol == i, rankInFiber == k.  Behind one NVSwitch every peer is equidistant, so the
// adjacency only fixes the LAYOUT (who owns which block), not performance.
// The three slice communicators of the reference (:86-88) are created on first use.
#pragma once
#include <cassert>
#include <iostream>
#include <memory>
#include <vector>

#include "hnh/comm.h"
#include "hnh/runtime.h"

class FlexibleGrid {
public:
    int i, j, k;
    int adjacency;
    int global_rank, num_procs;
    int dim_list[3];
    int nr, nc, nh;
    int permutation[3];

    std::shared_ptr<hnh::Comm> world;
    std::shared_ptr<hnh::Comm> row_world, col_world, fiber_world;
    int rankInRow, rankInCol, rankInFiber;

    FlexibleGrid(int nr, int nc, int nh, int adjacency) {
        world = hnh::Comm::world();
        num_procs = world->size();
        global_rank = world->rank();
        if (nr * nc * nh != num_procs)
            throw hnh::Error(-1, "FlexibleGrid: nr*nc*nh != number of ranks");
        if (adjacency < 1 || adjacency > 6) throw hnh::Error(-1, "FlexibleGrid: adjacency must be 1..6");
        dim_list[0] = this->nr = nr;
        dim_list[1] = this->nc = nc;
        dim_list[2] = this->nh = nh;
        this->adjacency = adjacency;
        // fastest-varying dimension first: 1 crf, 2 cfr, 3 rcf, 4 rfc, 5 fcr, 6 frc
        static const int perms[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
        for (int t = 0; t < 3; t++) permutation[t] = perms[adjacency - 1][t];

        get_ijk_indices(&i, &j, &k);
        assert(global_rank == get_global_rank(i, j, k));

        row_world = world->split(i + k * nr, j);
        col_world = world->split(j + k * nc, i);
        fiber_world = world->split(i + j * nr, k);
        rankInRow = row_world->rank();
        rankInCol = col_world->rank();
        rankInFiber = fiber_world->rank();
    }

    void get_ijk_indices(int rank, int *i, int *j, int *k) const {
        int t[3];
        t[permutation[0]] = rank % dim_list[permutation[0]];
        t[permutation[1]] = (rank / dim_list[permutation[0]]) % dim_list[permutation[1]];
        t[permutation[2]] = (rank / (dim_list[permutation[0]] * dim_list[permutation[1]])) % dim_list[permutation[2]];
        *i = t[0];
        *j = t[1];
        *k = t[2];
    }
    void get_ijk_indices(int *i, int *j, int *k) const { get_ijk_indices(global_rank, i, j, k); }

    int get_global_rank(int i, int j, int k) const {
        const int t[3] = {i, j, k};
        return t[permutation[0]] + t[permutation[1]] * dim_list[permutation[0]] +
               t[permutation[2]] * dim_list[permutation[0]] * dim_list[permutation[1]];
    }

    // slices (reference :86-88), created lazily: fixed k / fixed j / fixed i
    std::shared_ptr<hnh::Comm> rowcol_slice() {
        if (!rowcol_) rowcol_ = world->split(k, i + j * nr);
        return rowcol_;
    }
    std::shared_ptr<hnh::Comm> rowfiber_slice() {
        if (!rowfiber_) rowfiber_ = world->split(j, i + k * nr);
        return rowfiber_;
    }
    std::shared_ptr<hnh::Comm> colfiber_slice() {
        if (!colfiber_) colfiber_ = world->split(i, j + k * nc);
        return colfiber_;
    }

    void print_rank_information() const {
        std::cout << "Global Rank: " << global_rank << "i, j, k: (" << i << ", " << j << ", " << k << ")" << std::endl;
    }

    // Table of one int per rank, gathered on every rank, laid out [k][i][j] (self_test, :169-201).
    std::vector<int> gather_table(int msg) {
        std::vector<int> all(num_procs);
        world->host_allgather(&msg, all.data(), sizeof(int));
        std::vector<int> table(num_procs);
        for (int kk = 0; kk < nh; kk++)
            for (int ii = 0; ii < nr; ii++)
                for (int jj = 0; jj < nc; jj++) table[(kk * nr + ii) * nc + jj] = all[get_global_rank(ii, jj, kk)];
        return table;
    }

private:
    std::shared_ptr<hnh::Comm> rowcol_, rowfiber_, colfiber_;
};
