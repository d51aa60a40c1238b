/*
 * Cattle.hpp -- umbrella header of the B200-native C-ATTL3 hot path.
 *
 * Usage: put this directory BEFORE the reference's `C-ATTL3/` directory on the include path
 *
 *     g++ -std=c++11 -I<repo>/c-attl3_b200/cattle -I<repo>/include -I<reference>/C-ATTL3 -I<reference>/Eigen \
 *         app.cpp -L<repo>/c-attl3_b200 -lcattl3_b200
 *
 * and keep `#include "Cattle.hpp"` in the application as it is.  The headers below replace the
 * reference's kernel, activation, pooling and batch-normalisation layers, its FeedforwardNeuralNetwork /
 * ResidualNeuralNetwork layer loops and its SGD-family optimizers with device-resident implementations
 * that have the same class names and constructor signatures; each of them defines the include guard of
 * the reference header it stands in for, so when the reference's own umbrella header is pulled in at
 * the end (everything that is not on the hot path: losses, data providers, initialisations,
 * regularisations, the remaining layers and networks, GradientCheck), the originals of the replaced
 * headers are skipped.  Nothing of the reference is copied here: its headers are used where they lie.
 */
#ifndef C_ATTL3_B200_CATTLE_H_
#define C_ATTL3_B200_CATTLE_H_

// device runtime
#include "b200/Runtime.hpp"
#include "b200/Communicator.hpp"
#include "b200/DeviceLayer.hpp"
#include "b200/DeviceNetwork.hpp"
#include "b200/DeviceSequenceNetwork.hpp"
#include "b200/DeviceLoss.hpp"
#include "parameters/B200Parameters.hpp"

// the hot path
#include "layer/kernel/ConvKernelLayer.hpp"
#include "layer/kernel/TransConvKernelLayer.hpp"
#include "layer/kernel/DenseKernelLayer.hpp"
#include "layer/activation/ReLUActivationLayer.hpp"
#include "layer/activation/LeakyReLUActivationLayer.hpp"
#include "layer/activation/ELUActivationLayer.hpp"
#include "layer/activation/SwishActivationLayer.hpp"
#include "layer/activation/SigmoidActivationLayer.hpp"
#include "layer/activation/TanhActivationLayer.hpp"
#include "layer/activation/SoftplusActivationLayer.hpp"
#include "layer/activation/SoftmaxActivationLayer.hpp"
#include "layer/pool/MaxPoolLayer.hpp"
#include "layer/pool/MeanPoolLayer.hpp"
#include "layer/BatchNormLayer.hpp"
// the callers and neighbours of the hot path that keep a training step device resident
#include "layer/DropoutLayer.hpp"
#include "layer/ReshapeLayer.hpp"
#include "loss/SquaredLoss.hpp"
#include "loss/CrossEntropyLoss.hpp"
#include "neural_network/StackedNeuralNetwork.hpp"
#include "neural_network/DenseNeuralNetwork.hpp"
#include "neural_network/ParallelNeuralNetwork.hpp"
#include "neural_network/SequentialNeuralNetwork.hpp"
#include "neural_network/LSTMNeuralNetwork.hpp"
#include "neural_network/RecurrentNeuralNetwork.hpp"
#include "data_provider/MemoryDataProvider.hpp"
#include "neural_network/FeedforwardNeuralNetwork.hpp"
#include "neural_network/ResidualNeuralNetwork.hpp"
#include "optimizer/SGDOptimizer.hpp"
#include "optimizer/VanillaSGDOptimizer.hpp"
#include "optimizer/MomentumSGDOptimizer.hpp"
#include "optimizer/NesterovMomentumSGDOptimizer.hpp"
#include "optimizer/AdaGradOptimizer.hpp"
#include "optimizer/RMSPropOptimizer.hpp"
#include "optimizer/AdaDeltaOptimizer.hpp"
#include "optimizer/AdamOptimizer.hpp"
#include "optimizer/AdaMaxOptimizer.hpp"
#include "optimizer/NadamOptimizer.hpp"
#include "optimizer/AMSGradOptimizer.hpp"

// everything else, unchanged, from the reference (the next "Cattle.hpp" on the include path)
#include_next "Cattle.hpp"

#endif /* C_ATTL3_B200_CATTLE_H_ */
