/*
 * Compile-only check (tests/cpp/Makefile, `-fsyntax-only`): explicit instantiation of the B200 class templates in the
 * variants no test or example instantiates (stateful / multiplicatively integrating recurrent networks, every rank,
 * both scalars), so that every member of the header-only host side is at least type-checked against the reference's
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

} /* namespace cattle */
