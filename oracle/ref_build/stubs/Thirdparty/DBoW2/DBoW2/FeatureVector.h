// Thirdparty/DBoW2/DBoW2/FeatureVector.h — STAND-IN (the vendored header needs boost::serialization, which this image
// lacks).  Same container the reference iterates: std::map<NodeId, std::vector<unsigned int>> (FeatureVector.h:24-25).
#pragma once
#include <map>
#include <vector>
namespace DBoW2 {
typedef unsigned int NodeId;
class FeatureVector : public std::map<NodeId, std::vector<unsigned int>> {
public:
    void addFeature(NodeId id, unsigned int i_feature) { (*this)[id].push_back(i_feature); }
};
}  // namespace DBoW2
