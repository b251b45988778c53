#include "tf_syntax_stub.h"
