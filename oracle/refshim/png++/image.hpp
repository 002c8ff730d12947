/* refshim stand-in for png++ (TEST INFRASTRUCTURE; see ../refshim.h).
 * The reference uses png::image<png::gray_pixel_16> only in readPNG16 /
 * writePNG16 (adcensus.cu:1670-1705), which are outside the hot path; the
 * stub lets the file compile and fails loudly if those ops are called. */
#ifndef REFSHIM_PNGPP_IMAGE_HPP
#define REFSHIM_PNGPP_IMAGE_HPP
#include <stdint.h>
#include <stdexcept>
namespace png {
typedef uint16_t gray_pixel_16;
template <typename pixel>
class image {
public:
	explicit image(const char *) { throw std::runtime_error("refshim: PNG I/O is not available"); }
	image(size_t w, size_t h) : w_(w), h_(h) {}
	size_t get_width() const { return w_; }
	size_t get_height() const { return h_; }
	pixel get_pixel(size_t, size_t) const { return 0; }
	void set_pixel(size_t, size_t, pixel) {}
	void write(const char *) { throw std::runtime_error("refshim: PNG I/O is not available"); }
private:
	size_t w_ = 0, h_ = 0;
};
}
#endif
