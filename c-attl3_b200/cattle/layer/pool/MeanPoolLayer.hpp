/*
 * layer/pool/MeanPoolLayer.hpp -- B200 replacement of the reference's MeanPoolLayer
 * (C-ATTL3/layer/pool/MeanPoolLayer.hpp), same class template and constructors (2 x 2 windows with
 * stride 2 by default); defines the reference header's include guard.
 */
#ifndef C_ATTL3_LAYER_POOL_MEANPOOLLAYER_H_
#define C_ATTL3_LAYER_POOL_MEANPOOLLAYER_H_

#include "b200/SpatialPoolLayer.hpp"

namespace cattle {

/** Ranks 3 ({height, width, channels}) and 2 ({height, width}). */
template<typename Scalar, std::size_t Rank = 3>
class MeanPoolLayer : public b200::SpatialPoolLayer<Scalar,Rank,CATTL3_POOL_MEAN> {
	static_assert(Rank == 2 || Rank == 3, "rank 1 has its own specialisation");
	typedef Layer<Scalar,Rank> Root;
	typedef b200::SpatialPoolLayer<Scalar,Rank,CATTL3_POOL_MEAN> Core;
public:
	inline MeanPoolLayer(const typename Root::Dims& input_dims, std::size_t receptor_height = 2,
			std::size_t receptor_width = 2, std::size_t vertical_stride = 2, std::size_t horizontal_stride = 2) :
				Core(input_dims, receptor_height, receptor_width, vertical_stride, horizontal_stride) { }
	inline Root* clone() const {
		return new MeanPoolLayer(*this);
	}
};

/** Rank 1 ({length}). */
template<typename Scalar>
class MeanPoolLayer<Scalar,1> : public b200::SpatialPoolLayer<Scalar,1,CATTL3_POOL_MEAN> {
	typedef Layer<Scalar,1> Root;
	typedef b200::SpatialPoolLayer<Scalar,1,CATTL3_POOL_MEAN> Core;
public:
	inline MeanPoolLayer(const typename Root::Dims& input_dims, std::size_t receptor_length = 2,
			std::size_t stride = 2) :
				Core(input_dims, receptor_length, 1, stride, 1) { }
	inline Root* clone() const {
		return new MeanPoolLayer(*this);
	}
};

} /* namespace cattle */

#endif /* C_ATTL3_LAYER_POOL_MEANPOOLLAYER_H_ */
