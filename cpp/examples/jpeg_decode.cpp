// jpeg_decode: decode one baseline JPEG with the loader's own decoder (tnn/jpeg.hpp) and write the RGB bytes -- used by the tests to
// compare against an independent decoder, and handy for checking whether a file is one the C++ image loaders accept.
//   jpeg_decode <in.jpg> <out.rgb> [size]      prints "W H" (after the optional bilinear resize to size x size)
//   jpeg_decode --loader <root> <tiny_imagenet|imagenet100> <train|val> <batch> <out.f32>
//                                              first batch of the C++ image loader: prints "n classes cache label label ..." and
//                                              writes the batch (fp32 [B, S, S, 3])
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "tnn/data_loader.hpp"
#include "tnn/jpeg.hpp"

int main(int argc, char** argv) {
  if (argc >= 7 && std::string(argv[1]) == "--loader") {
    try {
      const bool train = std::string(argv[4]) == "train";
      std::unique_ptr<tnn::ImageFolderDataLoader> l;
      if (std::string(argv[3]) == "tiny_imagenet") l = std::make_unique<tnn::TinyImageNetDataLoader>(train);
      else l = std::make_unique<tnn::ImageNet100DataLoader>(train, 32);            // small crops keep the test quick
      if (!l->load_data(argv[2])) {
        std::cerr << "jpeg_decode: nothing to load under " << argv[2] << std::endl;
        return 1;
      }
      tnn::Tensor x, y;
      l->get_batch((size_t)atoi(argv[5]), x, y);
      std::cout << l->size() << " " << l->class_names().size() << " " << (l->uses_cache() ? 1 : 0);
      for (size_t i = 0; i < y->size(); ++i) std::cout << " " << y->data_as<int32_t>()[i];
      std::cout << std::endl;
      std::ofstream f(argv[6], std::ios::binary);
      f.write(static_cast<const char*>(x->data()), (std::streamsize)x->nbytes());
      return 0;
    } catch (const std::exception& e) {
      std::cerr << "jpeg_decode: " << e.what() << std::endl;
      return 1;
    }
  }
  if (argc < 3) {
    std::cerr << "usage: jpeg_decode <in.jpg> <out.rgb> [size]\n";
    return 2;
  }
  tnn::jpeg::Image im, rs;
  std::string err;
  if (!tnn::jpeg::decode_file(argv[1], &im, &err)) {
    std::cerr << "jpeg_decode: " << err << std::endl;
    return 1;
  }
  const tnn::jpeg::Image* out = &im;
  if (argc > 3) {
    const int s = atoi(argv[3]);
    tnn::jpeg::resize_bilinear(im, s, s, &rs);
    out = &rs;
  }
  std::ofstream f(argv[2], std::ios::binary);
  f.write(reinterpret_cast<const char*>(out->rgb.data()), (std::streamsize)out->rgb.size());
  std::cout << out->w << " " << out->h << std::endl;
  return 0;
}
