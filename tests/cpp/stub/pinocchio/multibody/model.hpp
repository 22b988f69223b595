// Minimal stand-in for the part of Pinocchio 3.0.0 that include/loik_b200/loik_pinocchio.hpp touches (TEST
// INFRASTRUCTURE: Pinocchio is not installed in the offline build container): pinocchio::Model with njoints / nv / nq /
// parents / jointPlacements / joints[i].{shortname, idx_q, idx_v, nq, nv}, SE3, Motion, Force and the aligned-vector macro.
// The joint axis of the *Unaligned types is a plain member here (LOIK_B200_PINOCCHIO_STUB); with the real library the
// adapter reads it through boost::get on the joint variant.
#pragma once
#define LOIK_B200_PINOCCHIO_STUB 1
#include <Eigen/Core>
#include <cstddef>
#include <string>
#include <vector>

#define PINOCCHIO_ALIGNED_STD_VECTOR(T) std::vector<T>

namespace pinocchio {
using Index = std::size_t;
using JointIndex = Index;

template <typename Scalar, int Options = 0>
struct SE3Tpl {
  using Matrix3 = Eigen::Matrix<Scalar, 3, 3>;
  using Vector3 = Eigen::Matrix<Scalar, 3, 1>;
  SE3Tpl() : rot(Matrix3::Identity()) {}
  SE3Tpl(const Matrix3& R, const Vector3& p) : rot(R), trans(p) {}
  static SE3Tpl Identity() { return SE3Tpl(); }
  const Matrix3& rotation() const { return rot; }
  const Vector3& translation() const { return trans; }
  Matrix3 rot;
  Vector3 trans;
};

template <typename Scalar, int Options = 0>
struct MotionTpl {  // [linear; angular]
  using Vector6 = Eigen::Matrix<Scalar, 6, 1>;
  MotionTpl() {}
  explicit MotionTpl(const Vector6& v) : v_(v) {}
  static MotionTpl Zero() { return MotionTpl(); }
  const Vector6& toVector() const { return v_; }
  Vector6 v_;
};
template <typename Scalar, int Options = 0>
struct ForceTpl {  // [linear; angular]
  using Vector6 = Eigen::Matrix<Scalar, 6, 1>;
  ForceTpl() {}
  explicit ForceTpl(const Vector6& v) : v_(v) {}
  static ForceTpl Zero() { return ForceTpl(); }
  const Vector6& toVector() const { return v_; }
  Vector6 v_;
};

struct JointModelStub {
  std::string name;  // "JointModelRZ", "JointModelRevoluteUnaligned", ...
  int iq = 0, iv = 0, nq_ = 1, nv_ = 1;
  Eigen::Matrix<double, 3, 1> axis;  // unaligned types
  std::string shortname() const { return name; }
  int idx_q() const { return iq; }
  int idx_v() const { return iv; }
  int nq() const { return nq_; }
  int nv() const { return nv_; }
};

template <typename Scalar, int Options = 0>
struct ModelTpl {
  using SE3 = SE3Tpl<Scalar, Options>;
  int njoints = 1, nv = 0, nq = 0;
  std::vector<JointIndex> parents{0};
  std::vector<SE3> jointPlacements{SE3()};
  std::vector<JointModelStub> joints{JointModelStub{"JointModelRZ"}};  // (entry 0: the universe)
  // test helper, not Pinocchio API: append a joint the way Model::addJoint does
  JointIndex addJoint(JointIndex parent, const std::string& shortname, const SE3& placement, int jnq = 1, int jnv = 1,
                      double ax = 0, double ay = 0, double az = 1) {
    JointModelStub j;
    j.name = shortname; j.iq = nq; j.iv = nv; j.nq_ = jnq; j.nv_ = jnv;
    j.axis[0] = ax; j.axis[1] = ay; j.axis[2] = az;
    joints.push_back(j); parents.push_back(parent); jointPlacements.push_back(placement);
    nq += jnq; nv += jnv;
    return (JointIndex)njoints++;
  }
};
using Model = ModelTpl<double>;
using SE3 = SE3Tpl<double>;
using Motion = MotionTpl<double>;
using Force = ForceTpl<double>;
}  // namespace pinocchio
