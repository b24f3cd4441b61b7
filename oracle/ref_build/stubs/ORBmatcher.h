// ORBmatcher.h — include shim used ONLY to build the reference's ORBmatcher.cc unmodified (oracle/ref_build/Makefile).
// The reference header includes MapPoint.h / KeyFrame.h / Frame.h, which need Eigen, PCL, boost, DBoW2 and g2o.  This
// file defines their include guards, provides stand-ins with the members ORBmatcher.cc reads (ref_types.h) and then
// hands over to the REAL header, /root/reference/orb_slam3/include/ORBmatcher.h, through #include_next.
#pragma once
#define MAPPOINT_H
#define KEYFRAME_H
#define FRAME_H
#include "ref_types.h"
#include_next "ORBmatcher.h"
