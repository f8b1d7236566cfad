// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for <JuceHeader.h>.
//
// The reference's hot-path translation units (MKLNonUniformConvolver.cpp, FFTBackend.cpp,
// eqprocessor/EQProcessor.{Processing,ProcessingCache,Coefficients}.cpp ...) are compiled *in place*
// from $CONVOPEQ_REF/src against this header so that the unmodified reference algorithm can be
// run as the parity checker.  Nothing here is reference code: these are our own definitions of the
// handful of JUCE names those TUs mention, written from JUCE's documented semantics.
// Release semantics: JUCE_DEBUG / _DEBUG / CONVOPEQ_ENABLE_RUNTIME_DIAGNOSTICS stay undefined.
#pragma once

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <vector>
#include <xmmintrin.h>
#include <pmmintrin.h>

#define JUCE_DECLARE_NON_COPYABLE(cls) \
    cls(const cls&) = delete;          \
    cls& operator=(const cls&) = delete;
#define JUCE_DECLARE_NON_COPYABLE_WITH_LEAK_DETECTOR(cls) JUCE_DECLARE_NON_COPYABLE(cls)
#define JUCE_LEAK_DETECTOR(cls)
#define jassert(x) ((void)0)
#define jassertfalse ((void)0)
#define DBG(x) ((void)0)
#define JUCE_ASSERT_MESSAGE_THREAD

namespace juce
{
using int64 = long long;
using uint64 = unsigned long long;
using uint32 = unsigned int;

template <typename T>
struct MathConstants
{
    static constexpr T pi = static_cast<T>(3.141592653589793238L);
    static constexpr T twoPi = static_cast<T>(2 * 3.141592653589793238L);
    static constexpr T halfPi = static_cast<T>(3.141592653589793238L / 2);
    static constexpr T euler = static_cast<T>(2.71828182845904523536L);
    static constexpr T sqrt2 = static_cast<T>(1.4142135623730950488L);
};

template <typename T>
constexpr T jmax(T a, T b) { return a < b ? b : a; }
template <typename T>
constexpr T jmax(T a, T b, T c) { return jmax(a, jmax(b, c)); }
template <typename T>
constexpr T jmin(T a, T b) { return b < a ? b : a; }
template <typename T>
constexpr T jmin(T a, T b, T c) { return jmin(a, jmin(b, c)); }
template <typename T>
constexpr T jlimit(T lo, T hi, T v) { return v < lo ? lo : (hi < v ? hi : v); }

template <typename... Ts>
inline void ignoreUnused(Ts&&...) noexcept {}

inline int nextPowerOfTwo(int n) noexcept
{
    --n;
    n |= (n >> 1);
    n |= (n >> 2);
    n |= (n >> 4);
    n |= (n >> 8);
    n |= (n >> 16);
    return n + 1;
}

inline bool isPowerOfTwo(int n) noexcept { return n > 0 && (n & (n - 1)) == 0; }

template <typename T>
inline T roundToInt(double v) noexcept { return static_cast<T>(std::lrint(v)); }
inline int roundToInt(double v) noexcept { return static_cast<int>(std::lrint(v)); }

struct FloatVectorOperations
{
    template <typename T> static void clear(T* d, size_t n) noexcept { std::memset(d, 0, n * sizeof(T)); }
    template <typename T> static void clear(T* d, int n) noexcept { if (n > 0) std::memset(d, 0, size_t(n) * sizeof(T)); }
    template <typename T> static void copy(T* d, const T* s, size_t n) noexcept { std::memcpy(d, s, n * sizeof(T)); }
    template <typename T> static void copy(T* d, const T* s, int n) noexcept { if (n > 0) std::memcpy(d, s, size_t(n) * sizeof(T)); }
    template <typename T> static void add(T* d, const T* s, int n) noexcept { for (int i = 0; i < n; ++i) d[i] += s[i]; }
    template <typename T> static void subtract(T* d, const T* s, int n) noexcept { for (int i = 0; i < n; ++i) d[i] -= s[i]; }
    template <typename T> static void multiply(T* d, T m, int n) noexcept { for (int i = 0; i < n; ++i) d[i] *= m; }
    template <typename T> static void multiply(T* d, const T* s, int n) noexcept { for (int i = 0; i < n; ++i) d[i] *= s[i]; }
    template <typename T> static void multiply(T* d, const T* s, T m, int n) noexcept { for (int i = 0; i < n; ++i) d[i] = s[i] * m; }
    template <typename T> static void fill(T* d, T v, int n) noexcept { for (int i = 0; i < n; ++i) d[i] = v; }
};

// FTZ + DAZ for the scope, like juce::ScopedNoDenormals on x86.
class ScopedNoDenormals
{
public:
    ScopedNoDenormals() noexcept : saved(_mm_getcsr()) { _mm_setcsr(saved | 0x8040u); }
    ~ScopedNoDenormals() noexcept { _mm_setcsr(saved); }
private:
    unsigned int saved;
};

class String
{
public:
    String() = default;
    String(const char* s) : str(s ? s : "") {}
    String(const std::string& s) : str(s) {}
    String(int v) : str(std::to_string(v)) {}
    String(unsigned v) : str(std::to_string(v)) {}
    String(long v) : str(std::to_string(v)) {}
    String(long long v) : str(std::to_string(v)) {}
    String(unsigned long v) : str(std::to_string(v)) {}
    String(unsigned long long v) : str(std::to_string(v)) {}
    String(double v) : str(std::to_string(v)) {}
    String(double v, int) : str(std::to_string(v)) {}
    String(float v) : str(std::to_string(v)) {}
    template <typename... A>
    static String formatted(const char* fmt, A... a)
    {
        char buf[2048];
        std::snprintf(buf, sizeof(buf), fmt, a...);
        return String(buf);
    }
    String operator+(const String& o) const { return String(str + o.str); }
    String& operator+=(const String& o) { str += o.str; return *this; }
    String& operator<<(const String& o) { str += o.str; return *this; }
    bool operator==(const String& o) const { return str == o.str; }
    bool operator!=(const String& o) const { return str != o.str; }
    bool isEmpty() const { return str.empty(); }
    bool isNotEmpty() const { return !str.empty(); }
    const char* toRawUTF8() const { return str.c_str(); }
    std::string toStdString() const { return str; }
    String trim() const { return *this; }
    int length() const { return (int) str.size(); }
    std::string str;
};
inline String operator+(const char* a, const String& b) { return String(a) + b; }

struct Logger
{
    static void writeToLog(const String&) {}
};

struct Decibels
{
    template <typename T>
    static T decibelsToGain(T dB, T minusInfinityDb = T(-100))
    {
        return dB > minusInfinityDb ? std::pow(T(10), dB * T(0.05)) : T();
    }
    template <typename T>
    static T gainToDecibels(T g, T minusInfinityDb = T(-100))
    {
        return g > T() ? jmax(minusInfinityDb, static_cast<T>(std::log10(g)) * T(20)) : minusInfinityDb;
    }
};

class Identifier
{
public:
    Identifier() = default;
    Identifier(const char* n) : name(n) {}
    Identifier(const String& n) : name(n) {}
    String toString() const { return name; }
    String name;
};

class var
{
public:
    var() = default;
    template <typename T> var(T) {}
    template <typename T> operator T() const { return T(); }
};

class File
{
public:
    File() = default;
    File(const String&) {}
    bool existsAsFile() const { return false; }
    String loadFileAsString() const { return {}; }
    String getFullPathName() const { return {}; }
};

class StringArray
{
public:
    int size() const { return 0; }
    String operator[](int) const { return {}; }
};

class ValueTree
{
public:
    ValueTree() = default;
    explicit ValueTree(const Identifier&) {}
    bool isValid() const { return false; }
    bool hasType(const Identifier&) const { return false; }
    bool hasProperty(const Identifier&) const { return false; }
    var getProperty(const Identifier&) const { return {}; }
    template <typename T> var getProperty(const Identifier&, T) const { return {}; }
    template <typename T> ValueTree& setProperty(const Identifier&, T, void*) { return *this; }
    void addChild(const ValueTree&, int, void*) {}
    void appendChild(const ValueTree&, void*) {}
    ValueTree getChildWithName(const Identifier&) const { return {}; }
    ValueTree getChild(int) const { return {}; }
    int getNumChildren() const { return 0; }
};

class ChangeListener;
class ChangeBroadcaster
{
public:
    virtual ~ChangeBroadcaster() = default;
    void sendChangeMessage() {}
    void addChangeListener(ChangeListener*) {}
    void removeChangeListener(ChangeListener*) {}
};

class MessageManager
{
public:
    static MessageManager* getInstance() { static MessageManager m; return &m; }
    static MessageManager* getInstanceWithoutCreating() { return nullptr; }
    bool isThisTheMessageThread() const { return false; }
    static void callAsync(std::function<void()>) {}
};

template <typename T>
class AudioBuffer
{
public:
    AudioBuffer() = default;
    AudioBuffer(T* const* chans, int nCh, int nS) : ch(chans, chans + nCh), ns(nS) {}
    int getNumChannels() const { return (int) ch.size(); }
    int getNumSamples() const { return ns; }
    const T* getReadPointer(int c) const { return ch[(size_t) c]; }
    const T* getReadPointer(int c, int o) const { return ch[(size_t) c] + o; }
    T* getWritePointer(int c) { return ch[(size_t) c]; }
    T* getWritePointer(int c, int o) { return ch[(size_t) c] + o; }
    void clear() { for (auto* p : ch) std::memset(p, 0, sizeof(T) * (size_t) ns); }
    void clear(int c, int s, int n) { std::memset(ch[(size_t) c] + s, 0, sizeof(T) * (size_t) n); }
    T getMagnitude(int c, int s, int n) const
    {
        T m {};
        for (int i = 0; i < n; ++i) m = jmax(m, (T) std::abs(ch[(size_t) c][s + i]));
        return m;
    }
    T getMagnitude(int s, int n) const
    {
        T m {};
        for (int c = 0; c < getNumChannels(); ++c) m = jmax(m, getMagnitude(c, s, n));
        return m;
    }
private:
    std::vector<T*> ch;
    int ns = 0;
};

template <typename T>
class SmoothedValue
{
public:
    SmoothedValue() = default;
    SmoothedValue(T v) : cur(v), tgt(v) {}
    void reset(double, double) {}
    void setCurrentAndTargetValue(T v) { cur = tgt = v; }
    void setTargetValue(T v) { tgt = v; cur = v; }
    T getNextValue() { return cur; }
    T getCurrentValue() const { return cur; }
    T getTargetValue() const { return tgt; }
    bool isSmoothing() const { return false; }
    void skip(int) {}
private:
    T cur {}, tgt {};
};

namespace dsp
{
template <typename T>
class AudioBlock
{
public:
    AudioBlock() = default;
    AudioBlock(T* const* chans, size_t nCh, size_t nS) : n(nS)
    {
        nch = nCh > 8 ? 8 : nCh;
        for (size_t i = 0; i < nch; ++i) ch[i] = chans[i];
    }
    AudioBlock(T* const* chans, size_t nCh, size_t start, size_t nS) : n(nS)
    {
        nch = nCh > 8 ? 8 : nCh;
        for (size_t i = 0; i < nch; ++i) ch[i] = chans[i] + start;
    }
    size_t getNumChannels() const noexcept { return nch; }
    size_t getNumSamples() const noexcept { return n; }
    T* getChannelPointer(size_t c) const noexcept { return ch[c]; }
    AudioBlock getSubBlock(size_t off, size_t len) const noexcept
    {
        AudioBlock b;
        b.nch = nch;
        b.n = len;
        for (size_t i = 0; i < nch; ++i) b.ch[i] = ch[i] + off;
        return b;
    }
    AudioBlock getSingleChannelBlock(size_t c) const noexcept
    {
        AudioBlock b;
        b.nch = 1;
        b.n = n;
        b.ch[0] = ch[c];
        return b;
    }
    void clear() const noexcept { for (size_t i = 0; i < nch; ++i) std::memset(ch[i], 0, sizeof(T) * n); }
private:
    T* ch[8] = {};
    size_t nch = 0;
    size_t n = 0;
};
} // namespace dsp
} // namespace juce
