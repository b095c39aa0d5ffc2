// qmcExceptions.h — same exception types and meanings as the reference's dmc/qmcExceptions.h:7-59,
// so messages stay familiar to its users.  The C ABI reports errors as strings; the host wrapper
// rethrows them as these types.
#pragma once
#include <exception>
#include <string>

#define QMC_EXCEPTION(name)                                            \
  class name : public std::exception {                                 \
   public:                                                             \
    explicit name(const std::string &m) : message(m) {}                \
    const char *what() const throw() override { return message.c_str(); } \
   private:                                                            \
    std::string message;                                               \
  }
QMC_EXCEPTION(missingImplementation);
QMC_EXCEPTION(invalidInput);
QMC_EXCEPTION(invalidState);
QMC_EXCEPTION(factoryIdNotRecorded);
#undef QMC_EXCEPTION
