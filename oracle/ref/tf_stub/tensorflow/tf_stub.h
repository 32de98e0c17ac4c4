#pragma once
/*
 * TEST INFRASTRUCTURE — minimal stand-in for the TensorFlow C++ headers named at
 * /root/reference/src/risk_game/player/alpha_zero/neural_network/alphazero_nn.h:13-25.
 * TensorFlow is not vendored by the reference and not installable here.  With
 * -D_DEBUG the reference compiles every session->Run out (alphazero_nn.cpp:174-178,
 * 245-264, 336-348) and only needs a Tensor that can be indexed, which is what lets
 * the reference's OWN input encoder (setInStateTensor, alphazero_nn.cpp:31-67) run.
 */
#include <string>
#include <vector>
#include <memory>
#include <initializer_list>
#include <cstdint>

namespace tensorflow
{
	typedef std::string tstring;
	enum DataType { DT_FLOAT = 1, DT_STRING = 7, DT_BOOL = 10 };

	struct TensorShape
	{
		std::vector<int64_t> dims;
		TensorShape() {}
		TensorShape(std::initializer_list<int64_t> d) : dims(d) {}
	};

	template <typename T> struct ScalarView
	{
		T* p;
		T& operator()() { return *p; }
	};

	template <typename T, int N> struct TensorView
	{
		T* p;
		int64_t d[N];
		int64_t dimension(int i) const { return d[i]; }
		T& operator()(int64_t a, int64_t b) const { return p[a * d[1] + b]; }
		T& operator()(int64_t a, int64_t b, int64_t c, int64_t e) const { return p[((a * d[1] + b) * d[2] + c) * d[3] + e]; }
	};

	class Tensor
	{
	public:
		DataType dtype = DT_FLOAT;
		TensorShape shape;
		std::shared_ptr<std::vector<float>> f;     /* DT_FLOAT payload */
		std::shared_ptr<std::vector<uint8_t>> b;   /* DT_BOOL payload  */
		std::shared_ptr<std::vector<tstring>> s;   /* DT_STRING payload */

		Tensor() {}
		Tensor(DataType t, const TensorShape& sh) : dtype(t), shape(sh)
		{
			int64_t n = 1;
			for (auto v : sh.dims) n *= v;
			f = std::make_shared<std::vector<float>>(n, 0.0f);
			b = std::make_shared<std::vector<uint8_t>>(n, 0);
			s = std::make_shared<std::vector<tstring>>(n);
		}

		template <typename T> ScalarView<T> scalar();
		template <typename T, int N> TensorView<T, N> tensor() const
		{
			TensorView<T, N> v;
			v.p = reinterpret_cast<T*>(f->data());
			for (int i = 0; i < N; i++) v.d[i] = shape.dims[i];
			return v;
		}
	};

	template <> inline ScalarView<float> Tensor::scalar<float>() { return ScalarView<float>{ f->data() }; }
	template <> inline ScalarView<bool> Tensor::scalar<bool>() { return ScalarView<bool>{ reinterpret_cast<bool*>(b->data()) }; }
	template <> inline ScalarView<tstring> Tensor::scalar<tstring>() { return ScalarView<tstring>{ s->data() }; }

	class GraphDef {};
	class Session { public: virtual ~Session() {} };
	namespace port { inline void InitMain(const char*, int*, char***) {} }
}
