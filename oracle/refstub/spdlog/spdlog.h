// TEST INFRASTRUCTURE ONLY -- logging shim for compiling the unmodified reference into oracle/_ref/
// (spdlog is an un-vendored dependency, gple/stdafx.h:53-55; the hot path only logs warnings).
#pragma once
#include <memory>
#include <string>
namespace spdlog
{
struct logger
{
};
template <typename... A>
inline void info(const A&...)
{
}
template <typename... A>
inline void warn(const A&...)
{
}
template <typename... A>
inline void error(const A&...)
{
}
template <typename... A>
inline void debug(const A&...)
{
}
inline void set_pattern(const std::string&) {}
inline void set_default_logger(std::shared_ptr<logger>) {}
inline std::shared_ptr<logger> stderr_color_mt(const std::string&)
{
	return std::make_shared<logger>();
}
} // namespace spdlog
