// Minimal stderr logger with the reference's macro surface (libs/logging/include/logging/log.hpp:7-11):
// LOG_WITH_LEVEL(level) << ...; levels ERROR / INFO / DEBUG; SET_LOG_LEVEL(level).
#pragma once
#include <iostream>
#include <sstream>

namespace logging {
enum LogLevel { ERROR = 0, INFO = 1, DEBUG = 2 };

inline LogLevel& reporting_level() {
    static LogLevel lvl = INFO;
    return lvl;
}

class Line {
   public:
    explicit Line(LogLevel l) : lvl_(l) {}
    ~Line() {
        static const char* tag[] = {"Error", "Info", "Debug"};
        if (lvl_ <= reporting_level()) std::cerr << tag[lvl_] << ": " << buf_.str() << std::endl;
    }
    std::ostringstream& stream() { return buf_; }

   private:
    LogLevel lvl_;
    std::ostringstream buf_;
};
}  // namespace logging

#define LOG_WITH_LEVEL(level) ::logging::Line(level).stream()
#define SET_LOG_LEVEL(level) (::logging::reporting_level() = (level))
