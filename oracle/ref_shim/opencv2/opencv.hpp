// oracle/ref_shim — TEST INFRASTRUCTURE: Library.cpp / SimpleMain.cpp include OpenCV (absent here) without the
// absolute-pose path using it; an empty namespace is all `using namespace cv;` needs.
namespace cv {}
