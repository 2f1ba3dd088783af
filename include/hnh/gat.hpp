// hnh/gat.hpp -- multi-head graph attention, forward pass only, on top of a Distributed_Sparse object:
// the second application that benchmark_algorithm can run (app == "gat").  Class and member names follow
// the reference (GATLayer, GAT{d_ops, layers, buffers, leaky_relu_alpha, computeSelfAttentionHead,
// forwardPass}: gat.hpp:26-113; used at benchmark_dist.cpp:88-94,133-135) so that its callers compile.
//
// One head of one layer:
//     H      = X_in * W                       dense projection          (hnh_dgemm_f64: own fp64 DMMA kernel)
//     e_uv   = <H_u, H_v> for every edge      SDDMM through d_ops       (K1)
//     e_uv   = LeakyReLU(e_uv)                                           (hnh_leaky_relu_f64)
//     Z      = E * H                          SpMM through d_ops        (K2)
//     X_out[:, head window] = ReLU(Z)                                    (hnh_relu_cols_f64)
// All operands stay in HBM between the steps.  The weights are whatever the caller stores in
// layers[i].wMats[j]; the reference leaves them zero and never assigns leaky_relu_alpha (reading it is
// undefined there) -- this implementation defaults alpha to 0.2.
#pragma once
#include <string>
#include <vector>

#include "hnh/distributed_sparse.h"

class GATLayer {
public:
    int input_features;
    int features_per_head;
    int num_heads;

    vector<DenseMatrix> wMats;  // one input_width x head_width projection per head
    VectorXd a1;                // attention vectors: declared by the reference, unused by its forward pass
    VectorXd a2;

    GATLayer(int in_width, int head_width, int heads)
        : input_features(in_width), features_per_head(head_width), num_heads(heads) {}
    int output_features() const { return features_per_head * num_heads; }
};

class GAT {
public:
    Distributed_Sparse *d_ops;
    vector<GATLayer> layers;
    vector<DenseMatrix> buffers;  // [0] network input (B-shaped), [i + 1] output of layer i (A-shaped)
    double leaky_relu_alpha = 0.2;

    GAT(vector<GATLayer> &l_input, Distributed_Sparse *ops) : d_ops(ops), layers(l_input) {
        if (layers.empty()) throw hnh::Error(-1, "GAT: no layers");
        for (size_t i = 1; i < layers.size(); i++)
            if (layers[i].input_features != layers[i - 1].output_features())
                throw hnh::Error(-1, "GAT: layer " + std::to_string(i) + " does not take the width layer " +
                                         std::to_string(i - 1) + " produces");
        // Shapes come from the algorithm object at the R in force, as in the reference (gat.hpp:60-79):
        // the input is B-shaped at R = input width, layer outputs are A-shaped at R = heads * head width,
        // a head's projection maps the previous buffer's local columns onto localAcols at R = head width.
        d_ops->setRValue(layers.front().input_features);
        buffers.push_back(d_ops->like_B_matrix(0.0));
        for (GATLayer &layer : layers) {
            const int64_t in_cols = buffers.back().cols();
            d_ops->setRValue(layer.output_features());
            buffers.push_back(d_ops->like_A_matrix(0.0));
            d_ops->setRValue(layer.features_per_head);
            layer.wMats.clear();
            for (int h = 0; h < layer.num_heads; h++)
                layer.wMats.push_back(DenseMatrix::Constant(in_cols, d_ops->localAcols, 0.0));
        }
    }

    // head j of layer i: reads buffers[i], writes its column window of buffers[i + 1]
    void computeSelfAttentionHead(int i, int j) {
        GATLayer &layer = layers.at((size_t)i);
        DenseMatrix &input = buffers[(size_t)i], &output = buffers[(size_t)i + 1];
        d_ops->setRValue(layer.features_per_head);

        DenseMatrix projected = input * layer.wMats.at((size_t)j);
        DenseMatrix neighbours = projected;  // the same projection in the role of the other SDDMM operand
        d_ops->de_shift(&neighbours, nullptr, k_spmmA);

        VectorXd pattern = d_ops->like_S_values(1.0);
        VectorXd scores = d_ops->like_S_values(1.0);
        d_ops->algorithm(projected, neighbours, pattern, &scores, k_sddmmA, true);
        scores.leakyRelu(leaky_relu_alpha);

        projected.setZero();  // becomes the SpMM accumulator
        d_ops->algorithm(projected, neighbours, scores, nullptr, k_spmmA, false);
        output.setMiddleColsRelu((int64_t)j * projected.cols(), projected);
    }

    void forwardPass() {
        for (size_t i = 0; i < layers.size(); i++)
            for (int j = 0; j < layers[i].num_heads; j++) computeSelfAttentionHead((int)i, j);
    }
};
