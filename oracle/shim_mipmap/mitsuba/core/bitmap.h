/* stand-in for <mitsuba/core/bitmap.h>: mipmap.h only NAMES Bitmap in the constructor that builds a pyramid from an image and in
 * toBitmap(); neither is instantiated by the shim (the pyramid comes in through the cache-file constructor). */
#pragma once
#include <mitsuba/mitsuba.h>
#include <mitsuba/core/spectrum.h> /* the reference's header: TSpectrum, Color3 */
namespace mitsuba {
class ReconstructionFilter;
class Bitmap : public Object {
public:
    enum EPixelFormat { ELuminance = 0, ELuminanceAlpha, ERGB, ERGBA };
    enum EComponentFormat { EUInt8 = 0, EFloat16, EFloat32, EFloat64, EFloat = EFloat32 };
    Bitmap(EPixelFormat, EComponentFormat, const Vector2i &);
    const Vector2i &getSize() const;
    int getWidth() const;
    int getHeight() const;
    size_t getPixelCount() const;
    Float getGamma() const;
    void *getData();
    ref<Bitmap> expand();
    ref<Bitmap> convert(EPixelFormat, EComponentFormat, Float, Float, Spectrum::EConversionIntent);
    ref<Bitmap> resample(const ReconstructionFilter *, int, int, const Vector2i &, Float, Float);
    template <typename T> static EComponentFormat componentFormat() { return EFloat32; }
};
}
