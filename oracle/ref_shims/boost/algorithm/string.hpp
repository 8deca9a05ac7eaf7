#pragma once
namespace boost {}
