/*
 * cattl3_oracle_impl.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's hot-path algorithms, instantiated twice by
 * cattl3_oracle.c (S = float with double accumulation, S = double).  Every function cites the
 * reference lines it restates (paths relative to /root/reference).  Layout everywhere is the
 * reference's: Eigen column-major tensors, i.e. offset(n,h,w,c) = n + N*(h + H*(w + W*c))
 * (C-ATTL3/core/EigenProxy.hpp:56-57) and column-major parameter matrices.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here against the compiled,
 * unmodified reference (oracle/_ref/libcattle_ref.so, built from /root/reference by
 * oracle/Makefile) and against the golden vectors in tests/golden/ that tests/golden/make_golden.py
 * generated from that same reference build.
 *
 * Macros expected: S (scalar type), FN(name) (symbol name mangler).
 */

#define IDX4(n, h, w, c, N, H, W) ((size_t)(n) + (size_t)(N) * ((size_t)(h) + (size_t)(H) * ((size_t)(w) + (size_t)(W) * (size_t)(c))))

/* Output spatial size: C-ATTL3/layer/kernel/ConvKernelLayer.hpp:194-197. */
static int FN(conv_out_dim)(int in, int r, int p, int d, int s) {
	return (in - r - (r - 1) * d + 2 * p) / s + 1;
}
/* C-ATTL3/layer/kernel/TransConvKernelLayer.hpp:200-203. */
static int FN(tconv_out_dim)(int in, int r, int p, int d, int s) {
	return (in - 1) * s + r + (r - 1) * d - 2 * p;
}

/*
 * ConvKernelLayerBase::_pass_forward, C-ATTL3/layer/kernel/ConvKernelLayer.hpp:115-148:
 * zero-pad, im2col with patch order (ow outer, oh inner, n fastest) and column order
 * k = rh + RH*(rw + RW*c), then cols * W + bias.  Restated as a direct sum.
 * ConvKernelLayerBase::_pass_back, :149-189: dW += cols^T dY, db += colsum(dY),
 * dX = crop(col2im(dY W^T)) unless input layer (dx == NULL).
 */
int FN(orc_conv)(const orc_geom* g, int transposed, const S* x, const S* w, const S* b, const S* dy,
		S* y, S* dx, S* dw, S* db, int back_reps, double* times_ms) {
	(void) times_ms;
	const int N = g->n, H = g->h, W = g->w, C = g->c, F = g->f, RH = g->rh, RW = g->rw;
	const int ph = g->ph, pw = g->pw, sh = g->sh, sw = g->sw, eh = g->dh + 1, ew = g->dw + 1;
	if (!transposed) {
		const int OH = FN(conv_out_dim)(H, RH, ph, g->dh, sh), OW = FN(conv_out_dim)(W, RW, pw, g->dw, sw);
		const size_t K = (size_t) RH * RW * C;
		if (y) {
			#pragma omp parallel for collapse(2) schedule(static)
			for (int f = 0; f < F; ++f)
			for (int ow = 0; ow < OW; ++ow)
			for (int oh = 0; oh < OH; ++oh)
			for (int n = 0; n < N; ++n) {
				double acc = 0;
				for (int c = 0; c < C; ++c)
				for (int rw = 0; rw < RW; ++rw) {
					int iw = ow * sw + rw * ew - pw;
					if (iw < 0 || iw >= W) continue;
					for (int rh = 0; rh < RH; ++rh) {
						int ih = oh * sh + rh * eh - ph;
						if (ih < 0 || ih >= H) continue;
						acc += (double) x[IDX4(n, ih, iw, c, N, H, W)] *
								(double) w[(size_t) rh + RH * ((size_t) rw + RW * (size_t) c) + K * f];
					}
				}
				y[IDX4(n, oh, ow, f, N, OH, OW)] = (S) (acc + (double) b[f]);
			}
		}
		if (!dy) return 0;
		/* weight + bias gradient */
		#pragma omp parallel for collapse(2) schedule(static)
		for (int f = 0; f < F; ++f)
		for (int c = 0; c < C; ++c)
		for (int rw = 0; rw < RW; ++rw)
		for (int rh = 0; rh < RH; ++rh) {
			double acc = 0;
			for (int ow = 0; ow < OW; ++ow) {
				int iw = ow * sw + rw * ew - pw;
				if (iw < 0 || iw >= W) continue;
				for (int oh = 0; oh < OH; ++oh) {
					int ih = oh * sh + rh * eh - ph;
					if (ih < 0 || ih >= H) continue;
					for (int n = 0; n < N; ++n)
						acc += (double) x[IDX4(n, ih, iw, c, N, H, W)] * (double) dy[IDX4(n, oh, ow, f, N, OH, OW)];
				}
			}
			dw[(size_t) rh + RH * ((size_t) rw + RW * (size_t) c) + K * f] = (S) (acc * back_reps);
		}
		for (int f = 0; f < F; ++f) {
			double acc = 0;
			for (size_t m = 0; m < (size_t) N * OH * OW; ++m)
				acc += (double) dy[m + (size_t) N * OH * OW * f];
			db[f] = (S) (acc * back_reps);
		}
		if (!dx) return 0;
		/* input gradient (gather form of the reference's col2im scatter-add) */
		#pragma omp parallel for collapse(2) schedule(static)
		for (int c = 0; c < C; ++c)
		for (int iw = 0; iw < W; ++iw)
		for (int ih = 0; ih < H; ++ih)
		for (int n = 0; n < N; ++n) {
			double acc = 0;
			for (int rw = 0; rw < RW; ++rw) {
				int tw = iw + pw - rw * ew;
				if (tw < 0 || tw % sw) continue;
				int ow = tw / sw;
				if (ow >= OW) continue;
				for (int rh = 0; rh < RH; ++rh) {
					int th = ih + ph - rh * eh;
					if (th < 0 || th % sh) continue;
					int oh = th / sh;
					if (oh >= OH) continue;
					for (int f = 0; f < F; ++f)
						acc += (double) dy[IDX4(n, oh, ow, f, N, OH, OW)] *
								(double) w[(size_t) rh + RH * ((size_t) rw + RW * (size_t) c) + K * f];
				}
			}
			dx[IDX4(n, ih, iw, c, N, H, W)] = (S) acc;
		}
		return 0;
	}
	/*
	 * TransConvKernelLayerBase::_pass_forward, C-ATTL3/layer/kernel/TransConvKernelLayer.hpp:115-151:
	 * (x as M x C) * W (C x RH*RW*F, col = rh + RH*(rw + RW*f)), col2im scatter-add into the padded
	 * output at (ih*sh + rh*(dh+1), iw*sw + rw*(dw+1)), crop, then a PER-OUTPUT-ELEMENT bias
	 * b[oh + OH*(ow + OW*f)].  _pass_back, :152-187: db += sum_n dY, G = im2col(pad(dY)),
	 * dW += x^T G, dX = G W^T.
	 */
	const int IH = H, IW = W;
	const int OH = FN(tconv_out_dim)(IH, RH, ph, g->dh, sh), OW = FN(tconv_out_dim)(IW, RW, pw, g->dw, sw);
	if (y) {
		#pragma omp parallel for collapse(2) schedule(static)
		for (int f = 0; f < F; ++f)
		for (int ow = 0; ow < OW; ++ow)
		for (int oh = 0; oh < OH; ++oh)
		for (int n = 0; n < N; ++n) {
			double acc = 0;
			for (int rw = 0; rw < RW; ++rw) {
				int tw = ow + pw - rw * ew;
				if (tw < 0 || tw % sw) continue;
				int iw = tw / sw;
				if (iw >= IW) continue;
				for (int rh = 0; rh < RH; ++rh) {
					int th = oh + ph - rh * eh;
					if (th < 0 || th % sh) continue;
					int ih = th / sh;
					if (ih >= IH) continue;
					for (int c = 0; c < C; ++c)
						acc += (double) x[IDX4(n, ih, iw, c, N, IH, IW)] *
								(double) w[(size_t) c + (size_t) C * ((size_t) rh + RH * ((size_t) rw + RW * (size_t) f))];
				}
			}
			y[IDX4(n, oh, ow, f, N, OH, OW)] = (S) (acc + (double) b[(size_t) oh + OH * ((size_t) ow + OW * (size_t) f)]);
		}
	}
	if (!dy) return 0;
	for (size_t e = 0; e < (size_t) OH * OW * F; ++e) {
		double acc = 0;
		for (int n = 0; n < N; ++n)
			acc += (double) dy[n + (size_t) N * e];
		db[e] = (S) (acc * back_reps);
	}
	#pragma omp parallel for collapse(2) schedule(static)
	for (int f = 0; f < F; ++f)
	for (int rw = 0; rw < RW; ++rw)
	for (int rh = 0; rh < RH; ++rh)
	for (int c = 0; c < C; ++c) {
		double acc = 0;
		for (int iw = 0; iw < IW; ++iw) {
			int ow = iw * sw + rw * ew - pw;
			if (ow < 0 || ow >= OW) continue;
			for (int ih = 0; ih < IH; ++ih) {
				int oh = ih * sh + rh * eh - ph;
				if (oh < 0 || oh >= OH) continue;
				for (int n = 0; n < N; ++n)
					acc += (double) x[IDX4(n, ih, iw, c, N, IH, IW)] * (double) dy[IDX4(n, oh, ow, f, N, OH, OW)];
			}
		}
		dw[(size_t) c + (size_t) C * ((size_t) rh + RH * ((size_t) rw + RW * (size_t) f))] = (S) (acc * back_reps);
	}
	if (!dx) return 0;
	#pragma omp parallel for collapse(2) schedule(static)
	for (int c = 0; c < C; ++c)
	for (int iw = 0; iw < IW; ++iw)
	for (int ih = 0; ih < IH; ++ih)
	for (int n = 0; n < N; ++n) {
		double acc = 0;
		for (int f = 0; f < F; ++f)
		for (int rw = 0; rw < RW; ++rw) {
			int ow = iw * sw + rw * ew - pw;
			if (ow < 0 || ow >= OW) continue;
			for (int rh = 0; rh < RH; ++rh) {
				int oh = ih * sh + rh * eh - ph;
				if (oh < 0 || oh >= OH) continue;
				acc += (double) dy[IDX4(n, oh, ow, f, N, OH, OW)] *
						(double) w[(size_t) c + (size_t) C * ((size_t) rh + RH * ((size_t) rw + RW * (size_t) f))];
			}
		}
		dx[IDX4(n, ih, iw, c, N, IH, IW)] = (S) acc;
	}
	return 0;
}

/*
 * DenseKernelLayer::pass_forward / pass_back, C-ATTL3/layer/kernel/DenseKernelLayer.hpp:92-115:
 * Y = X W + 1 b; dW += X^T dY; db += colsum(dY); dX = dY W^T.  X is n x in, column-major.
 */
int FN(orc_dense)(int n, int in, int out, const S* x, const S* w, const S* b, const S* dy, S* y, S* dx,
		S* dw, S* db, int back_reps, double* times_ms) {
	(void) times_ms;
	if (y) {
		#pragma omp parallel for schedule(static)
		for (int o = 0; o < out; ++o)
		for (int r = 0; r < n; ++r) {
			double acc = 0;
			for (int i = 0; i < in; ++i)
				acc += (double) x[r + (size_t) n * i] * (double) w[i + (size_t) in * o];
			y[r + (size_t) n * o] = (S) (acc + (double) b[o]);
		}
	}
	if (!dy) return 0;
	#pragma omp parallel for schedule(static)
	for (int o = 0; o < out; ++o) {
		double bacc = 0;
		for (int r = 0; r < n; ++r) bacc += (double) dy[r + (size_t) n * o];
		db[o] = (S) (bacc * back_reps);
		for (int i = 0; i < in; ++i) {
			double acc = 0;
			for (int r = 0; r < n; ++r)
				acc += (double) x[r + (size_t) n * i] * (double) dy[r + (size_t) n * o];
			dw[i + (size_t) in * o] = (S) (acc * back_reps);
		}
	}
	if (!dx) return 0;
	#pragma omp parallel for schedule(static)
	for (int i = 0; i < in; ++i)
	for (int r = 0; r < n; ++r) {
		double acc = 0;
		for (int o = 0; o < out; ++o)
			acc += (double) dy[r + (size_t) n * o] * (double) w[i + (size_t) in * o];
		dx[r + (size_t) n * i] = (S) acc;
	}
	return 0;
}

/*
 * Activation layers (kind numbering == CATTL3_ACT_*), x is n x vol column-major:
 *  0 ReLU      C-ATTL3/layer/activation/ReLUActivationLayer.hpp:45-57      y=max(x,0), x>=0 ? g : 0
 *  1 LeakyReLU C-ATTL3/layer/activation/LeakyReLUActivationLayer.hpp:50-62 y=max(x,a x), x>=0 ? g : a g
 *  2 ELU       C-ATTL3/layer/activation/ELUActivationLayer.hpp:55-78       x>=0 ? x : a(e^x-1); x>=0 ? g : (y+a) g
 *  3 Swish     C-ATTL3/layer/activation/SwishActivationLayer.hpp:45-61     s=1/(1+e^{-bx}); y=x s; s((1-s) b x+1) g
 *  4 Sigmoid   C-ATTL3/layer/activation/SigmoidActivationLayer.hpp         y=1/(1+e^-x); y(1-y) g
 *  5 Tanh      C-ATTL3/layer/activation/TanhActivationLayer.hpp            y=tanh x; (1-y^2) g
 *  6 Softplus  C-ATTL3/layer/activation/SoftplusActivationLayer.hpp:40-52  y=log(1+e^x); g/(1+e^-x)
 *  7 Softmax   C-ATTL3/layer/activation/SoftmaxActivationLayer.hpp:49-78   row-wise, max-subtracted,
 *              eps added to the denominator; dx_r = J_r^T g_r with J = diag(y) - y y^T
 */
int FN(orc_activation)(int kind, S alpha, int n, int vol, const S* x, const S* dy, S* y, S* dx) {
	size_t e = (size_t) n * vol;
	if (kind == 7) {
		const S eps = (S) 1e-5; /* NumericUtils<Scalar>::EPSILON2, C-ATTL3/core/NumericUtils.hpp:29 */
		for (int r = 0; r < n; ++r) {
			S mx = x[r];
			for (int j = 1; j < vol; ++j) if (x[r + (size_t) n * j] > mx) mx = x[r + (size_t) n * j];
			double sum = 0;
			for (int j = 0; j < vol; ++j) sum += exp((double) x[r + (size_t) n * j] - (double) mx);
			double* yr = (double*) malloc(sizeof(double) * vol);
			for (int j = 0; j < vol; ++j) {
				yr[j] = exp((double) x[r + (size_t) n * j] - (double) mx) / (sum + (double) eps);
				if (y) y[r + (size_t) n * j] = (S) yr[j];
			}
			if (dy && dx) {
				double dot = 0;
				for (int j = 0; j < vol; ++j) dot += yr[j] * (double) dy[r + (size_t) n * j];
				for (int j = 0; j < vol; ++j)
					dx[r + (size_t) n * j] = (S) (yr[j] * ((double) dy[r + (size_t) n * j] - dot));
			}
			free(yr);
		}
		return 0;
	}
	#pragma omp parallel for schedule(static)
	for (size_t i = 0; i < e; ++i) {
		double v = (double) x[i], a = (double) alpha, out, d;
		switch (kind) {
			case 0: out = v > 0 ? v : 0; d = v >= 0 ? 1 : 0; break;
			case 1: out = v > a * v ? v : a * v; d = v >= 0 ? 1 : a; break;
			case 2: out = v >= 0 ? v : a * (exp(v) - 1); d = v >= 0 ? 1 : out + a; break;
			case 3: { double s = 1 / (1 + exp(-a * v)); out = v * s; d = s * ((1 - s) * a * v + 1); break; }
			case 4: out = 1 / (1 + exp(-v)); d = out * (1 - out); break;
			case 5: out = tanh(v); d = 1 - out * out; break;
			case 6: out = log(1 + exp(v)); d = 1 / (1 + exp(-v)); break;
			default: out = v; d = 1;
		}
		if (y) y[i] = (S) out;
		if (dy && dx) dx[i] = (S) (d * (double) dy[i]);
	}
	return 0;
}

/*
 * PoolLayer::_pass_forward/_pass_back, C-ATTL3/layer/PoolLayer.hpp:77-116; no padding,
 * OH = (H-RH)/sh + 1 (:145-148).  kind 0 = MaxPoolLayerBase::_reduce/_d_reduce
 * (C-ATTL3/layer/pool/MaxPoolLayer.hpp:38-80): strict '>' starting from lowest(), scanned width-outer /
 * height-inner, so the FIRST maximum in (rw, rh) order wins; backward routes the gradient to that
 * element and overlapping windows accumulate.  kind 1 = MeanPoolLayerBase
 * (C-ATTL3/layer/pool/MeanPoolLayer.hpp:35-41): mean, backward g/(RH*RW) broadcast.
 */
int FN(orc_pool)(int kind, int n, int h, int w, int c, int rh, int rw, int sh, int sw, const S* x,
		const S* dy, S* y, S* dx, double* times_ms) {
	(void) times_ms;
	const int OH = (h - rh) / sh + 1, OW = (w - rw) / sw + 1;
	if (dy && dx) memset(dx, 0, sizeof(S) * (size_t) n * h * w * c);
	#pragma omp parallel for schedule(static)
	for (int ch = 0; ch < c; ++ch)
	for (int ow = 0; ow < OW; ++ow)
	for (int oh = 0; oh < OH; ++oh)
	for (int i = 0; i < n; ++i) {
		size_t o = IDX4(i, oh, ow, ch, n, OH, OW);
		if (kind == 0) {
			S best = -ORC_MAX; int bh = 0, bw = 0;
			for (int k = 0; k < rw; ++k)
			for (int l = 0; l < rh; ++l) {
				S v = x[IDX4(i, oh * sh + l, ow * sw + k, ch, n, h, w)];
				if (v > best) { best = v; bh = l; bw = k; }
			}
			if (y) y[o] = best;
			if (dy && dx) dx[IDX4(i, oh * sh + bh, ow * sw + bw, ch, n, h, w)] += dy[o];
		} else {
			double acc = 0;
			for (int k = 0; k < rw; ++k)
			for (int l = 0; l < rh; ++l)
				acc += (double) x[IDX4(i, oh * sh + l, ow * sw + k, ch, n, h, w)];
			if (y) y[o] = (S) (acc / (rh * rw));
			if (dy && dx) {
				S gv = dy[o] / (S) (rh * rw);
				for (int k = 0; k < rw; ++k)
				for (int l = 0; l < rh; ++l)
					dx[IDX4(i, oh * sh + l, ow * sw + k, ch, n, h, w)] += gv;
			}
		}
	}
	return 0;
}

/*
 * BatchNormLayer.  per_channel=1: BatchNormLayer<S,3,true>, C-ATTL3/layer/BatchNormLayer.hpp:225-262:
 * per channel over L = N*H*W elements: mu = mean, inv_sd = 1/sqrt(mean((x-mu)^2) + eps),
 * xhat = (x-mu) inv_sd, y = gamma xhat + beta; running stats: first batch assigns, later
 * (1-d) avg + d new (:234-243); inference uses (x - avg_mean) avg_inv_sd (:246).
 * backward (:250-261): dgamma += sum dy xhat, dbeta += sum dy,
 * dx = (L g - sum g - xhat sum(xhat g)) inv_sd / L with g = gamma dy.
 * per_channel=0: BatchNormLayer<S,3,false>, :337-391: same per activation (group = one of
 * H*W*C columns, L = N).
 * `steps` training passes over x[s], backward on the last, then inference on x[last].
 */
int FN(orc_batchnorm)(int per_channel, int n, int h, int w, int c, S decay, S eps, int steps, const S* x,
		const S* gamma, const S* beta, const S* dy, S* y, S* dx, S* dgamma, S* dbeta, S* run_mean,
		S* run_inv_sd, S* y_infer) {
	const size_t vol = (size_t) n * h * w * c;
	const size_t groups = per_channel ? (size_t) c : (size_t) h * w * c;
	const size_t L = per_channel ? (size_t) n * h * w : (size_t) n;
	double* rm = (double*) calloc(groups, sizeof(double));
	double* rs = (double*) calloc(groups, sizeof(double));
	double* mu = (double*) calloc(groups, sizeof(double));
	double* is = (double*) calloc(groups, sizeof(double));
	for (int s = 0; s < steps; ++s) {
		const S* xs = x + (size_t) s * vol;
		for (size_t gi = 0; gi < groups; ++gi) {
			const S* xg = xs + gi * L;
			double m = 0, v = 0;
			for (size_t i = 0; i < L; ++i) m += (double) xg[i];
			m /= (double) L;
			for (size_t i = 0; i < L; ++i) v += ((double) xg[i] - m) * ((double) xg[i] - m);
			v /= (double) L;
			mu[gi] = m;
			is[gi] = 1 / sqrt(v + (double) eps);
			if (s == 0) { rm[gi] = (double) (S) m; rs[gi] = (double) (S) is[gi]; }
			else {
				rm[gi] = (double) (S) ((1 - (double) decay) * rm[gi] + (double) decay * m);
				rs[gi] = (double) (S) ((1 - (double) decay) * rs[gi] + (double) decay * is[gi]);
			}
		}
	}
	const S* xs = x + (size_t) (steps - 1) * vol;
	for (size_t gi = 0; gi < groups; ++gi) {
		const S* xg = xs + gi * L;
		double gm = (double) gamma[gi], bt = (double) beta[gi];
		if (y) for (size_t i = 0; i < L; ++i)
			y[gi * L + i] = (S) (((double) xg[i] - mu[gi]) * is[gi] * gm + bt);
		if (y_infer) for (size_t i = 0; i < L; ++i)
			y_infer[gi * L + i] = (S) (((double) xg[i] - rm[gi]) * rs[gi] * gm + bt);
		if (run_mean) run_mean[gi] = (S) rm[gi];
		if (run_inv_sd) run_inv_sd[gi] = (S) rs[gi];
		if (dy) {
			const S* dg = dy + gi * L;
			double sg = 0, sxg = 0, sdy = 0, sdyx = 0;
			for (size_t i = 0; i < L; ++i) {
				double xh = ((double) xg[i] - mu[gi]) * is[gi];
				sdy += (double) dg[i];
				sdyx += (double) dg[i] * xh;
				sg += gm * (double) dg[i];
				sxg += xh * gm * (double) dg[i];
			}
			if (dgamma) dgamma[gi] = (S) sdyx;
			if (dbeta) dbeta[gi] = (S) sdy;
			if (dx) for (size_t i = 0; i < L; ++i) {
				double xh = ((double) xg[i] - mu[gi]) * is[gi];
				dx[gi * L + i] = (S) (((double) L * gm * (double) dg[i] - sg - xh * sxg) * is[gi] / (double) L);
			}
		}
	}
	free(rm); free(rs); free(mu); free(is);
	return 0;
}

/*
 * Optimizer update rules (kind numbering == CATTL3_OPT_*), hyper = {lr, a, b, eps}, restating
 * SGDOptimizer::_train's per-batch sequence regularize -> _update_params -> reset_grad
 * (C-ATTL3/optimizer/SGDOptimizer.hpp:57-70) for ONE parameter matrix with an optional L2 penalty
 * (grad += lambda * W, C-ATTL3/parameter_regularization/L2ParameterRegularization.hpp:31-33):
 *  0 VanillaSGD  VanillaSGDOptimizer.hpp:38-43          p -= lr g
 *  1 Momentum    MomentumSGDOptimizer.hpp:54-72         lr_e = lr/(1+a*epoch); v = b v + lr_e g; p -= v
 *  2 Nesterov    NesterovMomentumSGDOptimizer.hpp:43-55 v' = b v - lr_e g; p += -b v + (1+b) v'
 *  3 AdaGrad     AdaGradOptimizer.hpp:49-71             s += g^2; p -= lr g/(sqrt(s)+eps)
 *  4 RMSProp     RMSPropOptimizer.hpp:42-46             s = (1-b) s + b g^2; same step
 *  5 AdaDelta    AdaDeltaOptimizer.hpp:53-67            s=(1-a)s+a g^2; u=-g sqrt(d+eps)/sqrt(s+eps); p+=u; d=(1-a)d+a u^2
 *  6 Adam        AdamOptimizer.hpp:66-82                eps inside the bias corrections AND inside the sqrt
 *  7 AdaMax      AdaMaxOptimizer.hpp:44-60
 *  8 Nadam       NadamOptimizer.hpp:44-63
 *  9 AMSGrad     AMSGradOptimizer.hpp:50-67
 * The bias-correction scalars are evaluated in double and rounded to S exactly as the reference's
 * expression `(Scalar) 1 / (1 - pow(1 - l1_decay, timestep + 1) + epsilon)` does.
 */
int FN(orc_optimizer)(int kind, const S* hyper, S l2_lambda, int rows, int cols, int steps,
		int steps_per_epoch, const S* p0, const S* grads, S* p_out) {
	const size_t P = (size_t) rows * cols;
	const S lr = hyper[0], a = hyper[1], b = hyper[2], eps = hyper[3];
	S* s1 = (S*) calloc(P, sizeof(S));
	S* s2 = (S*) calloc(P, sizeof(S));
	S* s3 = (S*) calloc(P, sizeof(S));
	memcpy(p_out, p0, sizeof(S) * P);
	for (int t = 0; t < steps; ++t) {
		const S* gr = grads + (size_t) t * P;
		int epoch = t / steps_per_epoch;
		S lr_e = lr / (1 + a * epoch);
		S c1 = (S) ((S) 1 / (1 - pow(1 - a, t + 1) + eps));
		S c1n = (S) ((S) 1 / (1 - pow(1 - a, t + 2) + eps));
		S c2 = (S) ((S) 1 / (1 - pow(1 - b, t + 1) + eps));
		for (size_t i = 0; i < P; ++i) {
			S p = p_out[i];
			S g = gr[i] + (l2_lambda > 0 ? p * l2_lambda : 0);
			switch (kind) {
				case 0: p = p - g * lr; break;
				case 1: s1[i] = s1[i] * b + g * lr_e; p = p - s1[i]; break;
				case 2: { S old = s1[i]; s1[i] = old * b - g * lr_e; p = p + old * -b + s1[i] * (1 + b); break; }
				case 3: s1[i] += g * g; p = p - g * lr / ((S) sqrt(s1[i]) + eps); break;
				case 4: s1[i] = s1[i] * (1 - b) + g * g * b; p = p - g * lr / ((S) sqrt(s1[i]) + eps); break;
				case 5: {
					s1[i] = s1[i] * (1 - a) + g * g * a;
					S u = -g * (S) sqrt(s2[i] + eps) / (S) sqrt(s1[i] + eps);
					p = p + u;
					s2[i] = s2[i] * (1 - a) + u * u * a;
					break;
				}
				case 6:
					s1[i] = s1[i] * (1 - a) + g * a;
					s2[i] = s2[i] * (1 - b) + g * g * b;
					p = p - (s1[i] * (lr * c1)) / (S) sqrt(s2[i] * c2 + eps);
					break;
				case 7:
					s1[i] = s1[i] * (1 - a) + g * a;
					s2[i] = s2[i] * (1 - b) > (S) fabs(g) ? s2[i] * (1 - b) : (S) fabs(g);
					p = p - (s1[i] * (lr * c1)) / (s2[i] + eps);
					break;
				case 8:
					s1[i] = s1[i] * (1 - a) + g * a;
					s2[i] = s2[i] * (1 - b) + g * g * b;
					p = p - (g * (a * c1) + s1[i] * ((1 - a) * c1n)) * lr / (S) sqrt(s2[i] * c2 + eps);
					break;
				case 9:
					s1[i] = s1[i] * (1 - a) + g * a;
					s2[i] = s2[i] * (1 - b) + g * g * b;
					s3[i] = s2[i] > s3[i] ? s2[i] : s3[i];
					p = p - s1[i] * lr / (S) sqrt(s3[i] + eps);
					break;
				default: free(s1); free(s2); free(s3); return -1;
			}
			p_out[i] = p;
		}
	}
	free(s1); free(s2); free(s3);
	return 0;
}

#undef IDX4
