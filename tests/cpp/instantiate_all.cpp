/*
 * Compile-only check (tests/cpp/Makefile, `-fsyntax-only`): explicit instantiation of the B200 class templates in the
 * variants no test or example instantiates (stateful / multiplicatively integrating recurrent networks, the containers,
 * layers, losses, data provider and all ten optimizers at every rank, for sequential and non-sequential data, both
 * scalars), so that every member of the header-only host side is at least type-checked against the reference's
 * base classes.  (StackedNeuralNetwork is left to implicit instantiation: it holds members for the other value of its
 * Sequential parameter that are, by design, never instantiated.)
 */
#include "Cattle.hpp"

namespace cattle {

#define INSTANTIATE_SEQUENCE_NETS(S) \
	template class SequentialNeuralNetwork<S,1>; \
	template class SequentialNeuralNetwork<S,2>; \
	template class SequentialNeuralNetwork<S,3>; \
	template class LSTMNeuralNetwork<S,1,false,true>; \
	template class LSTMNeuralNetwork<S,2,true,false>; \
	template class LSTMNeuralNetwork<S,3,true,true>; \
	template class RecurrentNeuralNetwork<S,1,false,true>; \
	template class RecurrentNeuralNetwork<S,2,true,false>; \
	template class RecurrentNeuralNetwork<S,3,true,true>; \
	template class NadamOptimizer<S,1,true>; \
	template class NadamOptimizer<S,2,true>; \
	template class MomentumSGDOptimizer<S,3,true>;

INSTANTIATE_SEQUENCE_NETS(float)
INSTANTIATE_SEQUENCE_NETS(double)

#define INSTANTIATE_CONTAINERS(S, R) \
	template class FeedforwardNeuralNetwork<S,R>; \
	template class ResidualNeuralNetwork<S,R>; \
	template class ParallelNeuralNetwork<S,R,PARALLEL_CONCAT_LO_RANK>; \
	template class ParallelNeuralNetwork<S,R,PARALLEL_CONCAT_HI_RANK>; \
	template class ParallelNeuralNetwork<S,R,PARALLEL_SUM>; \
	template class ParallelNeuralNetwork<S,R,PARALLEL_MUL>; \
	template class DenseNeuralNetwork<S,R,DENSE_LOWEST_RANK>; \
	template class DenseNeuralNetwork<S,R,DENSE_HIGHEST_RANK>; \
	template class DropoutLayer<S,R>; \
	template class ReshapeLayer<S,R>; \
	template class DenseKernelLayer<S,R>;

INSTANTIATE_CONTAINERS(float, 1)
INSTANTIATE_CONTAINERS(float, 2)
INSTANTIATE_CONTAINERS(float, 3)
INSTANTIATE_CONTAINERS(double, 1)
INSTANTIATE_CONTAINERS(double, 2)
INSTANTIATE_CONTAINERS(double, 3)

#define INSTANTIATE_OPTIMIZERS(S, SEQ) \
	template class VanillaSGDOptimizer<S,3,SEQ>; \
	template class MomentumSGDOptimizer<S,2,SEQ>; \
	template class NesterovMomentumSGDOptimizer<S,3,SEQ>; \
	template class AdaGradOptimizer<S,3,SEQ>; \
	template class RMSPropOptimizer<S,3,SEQ>; \
	template class AdaDeltaOptimizer<S,3,SEQ>; \
	template class AdamOptimizer<S,3,SEQ>; \
	template class AdaMaxOptimizer<S,3,SEQ>; \
	template class NadamOptimizer<S,3,SEQ>; \
	template class AMSGradOptimizer<S,3,SEQ>; \
	template class SquaredLoss<S,3,SEQ>; \
	template class CrossEntropyLoss<S,3,SEQ>; \
	template class MemoryDataProvider<S,3,SEQ,true>;

INSTANTIATE_OPTIMIZERS(float, false)
INSTANTIATE_OPTIMIZERS(float, true)
INSTANTIATE_OPTIMIZERS(double, false)
INSTANTIATE_OPTIMIZERS(double, true)

} /* namespace cattle */
