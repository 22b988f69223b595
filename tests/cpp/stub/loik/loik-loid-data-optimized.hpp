// Stand-in for loik::IkIdDataTypeOptimizedTpl (/root/reference/include/loik/loik-loid-data-optimized.hpp:62-379, ctor
// .hxx:40-104) with the members the adapter writes back and the reference's tests read (tests/loik-loid.cpp:597-615).
// TEST INFRASTRUCTURE: same member names, types and sizes as the reference's struct (indexed by joint id 0..njoints-1, or
// by constraint slot), on the stub Eigen / Pinocchio types.
#pragma once
#include <numeric>
#include <pinocchio/multibody/model.hpp>

namespace loik {
template <typename _Scalar, int _Options = 0>
struct IkIdDataTypeOptimizedTpl {
  typedef _Scalar Scalar;
  typedef pinocchio::ModelTpl<Scalar, _Options> Model;
  typedef pinocchio::SE3Tpl<Scalar, _Options> SE3;
  typedef pinocchio::MotionTpl<Scalar, _Options> Motion;
  typedef pinocchio::ForceTpl<Scalar, _Options> Force;
  typedef pinocchio::Index Index;
  typedef std::vector<Index> IndexVector;
  typedef Eigen::Matrix<Scalar, Eigen::Dynamic, 1, _Options> DVec;
  typedef Eigen::Matrix<Scalar, 6, 1, _Options> Vec6;
  typedef Eigen::Matrix<Scalar, 6, 6> Mat6x6;

  IkIdDataTypeOptimizedTpl(const Model& model, const int num_eq_c_)
      : liMi((size_t)model.njoints), nu(model.nv), vis((size_t)model.njoints), His((size_t)model.njoints, Mat6x6::Identity()),
        pis((size_t)model.njoints), r(model.nv), fis((size_t)model.njoints), yis((size_t)num_eq_c_), w(model.nv), z(model.nv),
        num_eq_c(num_eq_c_), Aty((size_t)num_eq_c_), fis_diff_plus_Aty((size_t)model.njoints), Stf_plus_w(model.nv),
        joint_full_range((size_t)model.njoints), joint_range((size_t)model.njoints - 1) {
    std::iota(joint_full_range.begin(), joint_full_range.end(), 0);
    std::iota(joint_range.begin(), joint_range.end(), 1);
  }
  PINOCCHIO_ALIGNED_STD_VECTOR(SE3) liMi;
  DVec nu;
  PINOCCHIO_ALIGNED_STD_VECTOR(Motion) vis;
  PINOCCHIO_ALIGNED_STD_VECTOR(Mat6x6) His;
  PINOCCHIO_ALIGNED_STD_VECTOR(Force) pis;
  DVec r;
  PINOCCHIO_ALIGNED_STD_VECTOR(Force) fis;
  PINOCCHIO_ALIGNED_STD_VECTOR(Vec6) yis;
  DVec w, z;
  int num_eq_c;
  PINOCCHIO_ALIGNED_STD_VECTOR(Vec6) Aty;
  PINOCCHIO_ALIGNED_STD_VECTOR(Force) fis_diff_plus_Aty;
  DVec Stf_plus_w;
  IndexVector joint_full_range, joint_range;
};
typedef IkIdDataTypeOptimizedTpl<double> IkIdDataOptimized;
}  // namespace loik
