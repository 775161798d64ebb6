// Stand-in for <grid_map_core/GridMap.hpp>: just enough of grid_map::GridMap for the
// plugin shells and their test harness.  Layers are column-major float32 like Eigen::MatrixXf.
#pragma once
#include <cmath>
#include <limits>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>
namespace grid_map {
struct Index2 {
  int v[2];
  int operator()(int k) const { return v[k]; }
  int& operator()(int k) { return v[k]; }
};
struct Vec2 {
  double v[2];
  double operator()(int k) const { return v[k]; }
  double& operator()(int k) { return v[k]; }
  double x() const { return v[0]; }
  double y() const { return v[1]; }
};
typedef Index2 Size;
typedef Index2 Index;
typedef Vec2 Length;
typedef Vec2 Position;
class Matrix {
 public:
  Matrix() {}
  Matrix(int r, int c, float fill) : rows_(r), cols_(c), d_((size_t)r * c, fill) {}
  float* data() { return d_.data(); }
  const float* data() const { return d_.data(); }
  int rows() const { return rows_; }
  int cols() const { return cols_; }
  float& operator()(int i, int j) { return d_[(size_t)j * rows_ + i]; }
  float operator()(int i, int j) const { return d_[(size_t)j * rows_ + i]; }

 private:
  int rows_ = 0, cols_ = 0;
  std::vector<float> d_;
};
class GridMap {
 public:
  void setGeometry(const Length& length, double resolution, const Position& position) {
    size_(0) = (int)std::round(length(0) / resolution);
    size_(1) = (int)std::round(length(1) / resolution);
    resolution_ = resolution;
    length_(0) = size_(0) * resolution;
    length_(1) = size_(1) * resolution;
    position_ = position;
    start_(0) = start_(1) = 0;
  }
  void add(const std::string& layer, float value = std::numeric_limits<float>::quiet_NaN()) { data_[layer] = Matrix(size_(0), size_(1), value); }
  bool erase(const std::string& layer) { return data_.erase(layer) > 0; }
  bool exists(const std::string& layer) const { return data_.count(layer) > 0; }
  const Matrix& get(const std::string& layer) const {
    auto it = data_.find(layer);
    if (it == data_.end()) throw std::out_of_range("GridMap::get(...) : No map layer '" + layer + "' available.");
    return it->second;
  }
  Matrix& get(const std::string& layer) {
    auto it = data_.find(layer);
    if (it == data_.end()) throw std::out_of_range("GridMap::get(...) : No map layer '" + layer + "' available.");
    return it->second;
  }
  const Matrix& operator[](const std::string& layer) const { return get(layer); }
  Matrix& operator[](const std::string& layer) { return get(layer); }
  const Size& getSize() const { return size_; }
  double getResolution() const { return resolution_; }
  const Length& getLength() const { return length_; }
  const Position& getPosition() const { return position_; }
  const Index& getStartIndex() const { return start_; }
  void setStartIndex(const Index& s) { start_ = s; }
  void convertToDefaultStartIndex() {  // unwrap the circular buffer (rows and columns rotate independently)
    if (start_(0) == 0 && start_(1) == 0) return;
    for (auto& kv : data_) {
      Matrix out(size_(0), size_(1), 0.f);
      for (int j = 0; j < size_(1); ++j)
        for (int i = 0; i < size_(0); ++i) out(i, j) = kv.second((i + start_(0)) % size_(0), (j + start_(1)) % size_(1));
      kv.second = out;
    }
    start_(0) = start_(1) = 0;
  }
  std::vector<std::string> getLayers() const {
    std::vector<std::string> l;
    for (auto& kv : data_) l.push_back(kv.first);
    return l;
  }

 private:
  std::unordered_map<std::string, Matrix> data_;
  Size size_{{0, 0}};
  double resolution_ = 0.0;
  Length length_{{0, 0}};
  Position position_{{0, 0}};
  Index start_{{0, 0}};
};
}  // namespace grid_map
